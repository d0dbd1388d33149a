"""tfkaldi_b200 — B200-native engine for the DNN / CrossEnthropyTrainer / Decoder hot path of
vrenkens/tfkaldi, behind the reference's own Python API (neuralNetworks.*, processing.*).

Arithmetic lives in libtfkaldi_b200.so (hand-written sm_100a CUDA behind a C-ABI, see
include/tfkaldi_b200.h); this package is the host-side mirror of the reference interface.
"""
__version__ = "0.1.0"
