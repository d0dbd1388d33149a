"""Classifier contract (reference: neuralNetworks/classifiers/classifier.py:16-37)."""
from abc import ABCMeta, abstractmethod


class Classifier(object, metaclass=ABCMeta):
    """a classifier maps packed input frames to output logits"""

    def __init__(self, output_dim):
        self.output_dim = output_dim

    @abstractmethod
    def engine_spec(self, input_dim):
        """description of the network for tfkaldi_b200.engine.Engine (replaces graph construction in
        Classifier.__call__(inputs, seq_length, is_training, reuse, scope))"""
        raise NotImplementedError("Abstract method")
