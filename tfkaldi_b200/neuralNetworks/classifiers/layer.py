"""Feed-forward layer specification (reference: neuralNetworks/classifiers/layer.py)."""
import numpy as np


class FFLayer(object):
    """y = activation(x W + b); W [in, out] row-major, b [out]   (layer.py:24-58)"""

    def __init__(self, output_dim, activation, weights_std=None):
        self.output_dim = output_dim
        self.activation = activation
        self.weights_std = weights_std

    def initial_parameters(self, input_dim, rng):
        """W ~ N(0, std^2) with std = weights_std if given (0 is 'given': the output layer) else
        1/sqrt(input_dim); b = 0   (layer.py:39-48)"""
        std = self.weights_std if self.weights_std is not None else 1.0 / input_dim ** 0.5
        weights = (rng.standard_normal((input_dim, self.output_dim)) * std).astype(np.float32)
        return weights, np.zeros(self.output_dim, dtype=np.float32)
