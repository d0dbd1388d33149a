"""Classifier *specifications* (the reference builds TF graphs here; we describe the network and hand
the description to the CUDA engine)."""
from . import activation, classifier, dnn, layer  # noqa: F401
