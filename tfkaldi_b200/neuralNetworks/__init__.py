"""Host-side mirror of the reference's `neuralNetworks` package: same class names, constructor
arguments and call order; the TensorFlow graph/session is replaced by tfkaldi_b200.engine.Engine."""
from . import classifiers, decoder, nnet, trainer  # noqa: F401
