"""Target coders (reference: processing/target_coder.py).  Only the alignment flavour is on the hot
path; TextCoder belongs to the reference's unfinished CTC branch and is out of scope."""
from abc import ABCMeta, abstractmethod

import numpy as np


class TargetCoder(object, metaclass=ABCMeta):
    """maps space separated target strings to uint32 id vectors"""

    def __init__(self, target_normalizer):
        self.target_normalizer = target_normalizer
        self.lookup = {symbol: index for index, symbol in enumerate(self.create_alphabet())}

    @abstractmethod
    def create_alphabet(self):
        """list of all target symbols, position = id"""

    def encode(self, targets):
        """'p1 p2 ...' -> np.uint32 [len]   (target_coder.py:36-55); unknown symbols raise KeyError"""
        normalized = self.target_normalizer(targets, self.lookup.keys())
        lookup = self.lookup
        return np.array([lookup[symbol] for symbol in normalized.split(" ")], dtype=np.uint32)

    def decode(self, encoded_targets):
        symbols = list(self.lookup.keys())
        return " ".join(symbols[index] for index in encoded_targets)

    @property
    def num_labels(self):
        return len(self.lookup)


class AlignmentCoder(TargetCoder):
    """pdf-id alignments: alphabet '0' .. str(num_targets-1)   (target_coder.py:120-142)"""

    def __init__(self, target_normalizer, num_targets):
        self.num_targets = num_targets
        super(AlignmentCoder, self).__init__(target_normalizer)

    def create_alphabet(self):
        return [str(target) for target in range(self.num_targets)]
