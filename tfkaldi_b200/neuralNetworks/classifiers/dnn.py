"""The DNN classifier specification (reference: neuralNetworks/classifiers/dnn.py)."""
from .activation import TfActivation, compile_chain, linear
from .classifier import Classifier
from .layer import FFLayer


class DNN(Classifier):
    """num_layers hidden FFLayers of num_units with a shared activation chain, then a zero-initialised
    linear output layer of output_dim   (dnn.py:13-35, 64-68)"""

    def __init__(self, output_dim, num_layers, num_units, activation, layerwise_init=True):
        super(DNN, self).__init__(output_dim)
        self.num_layers = num_layers
        self.num_units = num_units
        self.activation = activation
        self.layerwise_init = layerwise_init

    def engine_spec(self, input_dim):
        spec = compile_chain(self.activation)
        spec.update(num_layers=self.num_layers, input_dim=input_dim, hidden_dim=self.num_units, output_dim=self.output_dim)
        return spec

    def initial_parameters(self, input_dim, rng):
        """{'W{l}', 'b{l}', ['beta{l}', 'moving_mean{l}', 'moving_var{l}']}: what init_op.run() would
        draw (trainer.py:244-247)."""
        import numpy as np

        hidden = FFLayer(self.num_units, self.activation)
        out = FFLayer(self.output_dim, TfActivation(None, linear), 0)  # dnn.py:67-68
        params = {}
        batch_norm = compile_chain(self.activation)["batch_norm"]
        for l in range(self.num_layers):
            params["W%d" % l], params["b%d" % l] = hidden.initial_parameters(input_dim if l == 0 else self.num_units, rng)
            if batch_norm:
                params["beta%d" % l] = np.zeros(self.num_units, np.float32)
                params["moving_mean%d" % l] = np.zeros(self.num_units, np.float32)
                params["moving_var%d" % l] = np.ones(self.num_units, np.float32)
        params["W%d" % self.num_layers], params["b%d" % self.num_layers] = out.initial_parameters(self.num_units, rng)
        return params
