"""Activation chain specifications (reference: neuralNetworks/classifiers/activation.py).

The reference composes decorators, each applying the wrapped activation first and its own function
after (activation.py:33-40); Nnet builds  Batchnorm -> nonlinearity -> L2Norm -> Dropout
(nnet.py:42-72).  Here the same constructors record the chain; `stages()` lists it inner-first and the
engine fuses it into the FFLayer kernel epilogue."""
from abc import ABCMeta, abstractmethod


def relu(x=None):
    """stands for tf.nn.relu"""
    return "relu"


def sigmoid(x=None):
    """stands for tf.nn.sigmoid"""
    return "sigmoid"


def tanh(x=None):
    """stands for tf.nn.tanh"""
    return "tanh"


def linear(x=None):
    """stands for `lambda(x): x` (nnet.py:62)"""
    return "linear"


_KNOWN = {"relu": "relu", "sigmoid": "sigmoid", "tanh": "tanh", "linear": "linear", "identity": "linear", "<lambda>": "linear"}


class Activation(object, metaclass=ABCMeta):
    def __init__(self, activation=None):
        self.activation = activation

    def stages(self):
        """[(kind, argument), ...] applied in order, wrapped activation first (activation.py:33-40)"""
        inner = self.activation.stages() if self.activation is not None else []
        return inner + [self._stage()]

    @abstractmethod
    def _stage(self):
        """this wrapper's own stage (the reference's _apply_func)"""
        raise NotImplementedError("Abstract method")


class TfActivation(Activation):
    """element-wise nonlinearity (activation.py:58-84).  `tfActivation` is one of this module's
    relu / sigmoid / tanh / linear stand-ins, or a name."""

    def __init__(self, activation, tfActivation):
        super(TfActivation, self).__init__(activation)
        self.tf_activation = tfActivation

    def _stage(self):
        name = self.tf_activation if isinstance(self.tf_activation, str) else getattr(self.tf_activation, "__name__", "")
        if name not in _KNOWN:
            raise Exception("unkown nonlinearity")  # nnet.py:65 (sic)
        return ("nonlin", _KNOWN[name])


class L2Norm(Activation):
    """per-frame mean-square normalisation (activation.py:87-111)"""

    def _stage(self):
        return ("l2norm", None)


class Dropout(Activation):
    """dropout with KEEP probability `dropout` in (0, 1] (activation.py:113-143)"""

    def __init__(self, activation, dropout):
        super(Dropout, self).__init__(activation)
        assert dropout > 0 and dropout <= 1
        self.dropout = dropout

    def _stage(self):
        return ("dropout", float(self.dropout))


class Batchnorm(Activation):
    """tf.contrib.layers.batch_norm with its defaults (activation.py:145-161)"""

    def _stage(self):
        return ("batchnorm", None)


def compile_chain(activation):
    """Activation chain -> engine keyword arguments.  The engine fuses exactly the order Nnet builds
    (nnet.py:42-72); anything else is refused loudly instead of being silently re-ordered."""
    stages = activation.stages() if activation is not None else []
    spec = {"batch_norm": False, "nonlin": "linear", "keep_prob": 1.0, "l2_norm": False}
    order = {"batchnorm": 0, "nonlin": 1, "l2norm": 2, "dropout": 3}
    last = -1
    for kind, arg in stages:
        if order[kind] <= last:
            raise NotImplementedError("activation chain %r is not in the order batchnorm -> nonlinearity -> l2norm -> dropout" % (stages,))
        last = order[kind]
        if kind == "batchnorm":
            spec["batch_norm"] = True
        elif kind == "nonlin":
            spec["nonlin"] = arg  # relu / sigmoid / tanh / linear (nnet.py:47-62)
        elif kind == "l2norm":
            spec["l2_norm"] = True
        elif kind == "dropout":
            spec["keep_prob"] = arg
    return spec
