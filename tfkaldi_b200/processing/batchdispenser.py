"""Utterance batch dispensers (reference: processing/batchdispenser.py).  The alignment flavour
feeds the trainer; the text flavour belongs to the unfinished CTC branch and is not built."""
import gzip
from abc import ABCMeta, abstractmethod

import numpy as np


class BatchDispenser(object, metaclass=ABCMeta):
    """dispenses `size` utterances (spliced features + encoded targets) per call"""

    @abstractmethod
    def read_target_file(self, target_path):
        """-> {utterance id: target string}"""

    def __init__(self, feature_reader, target_coder, size, target_path):
        self.feature_reader = feature_reader
        self.target_dict = self.read_target_file(target_path)
        self.size = size
        self.target_coder = target_coder
        # The reference encodes every target string here (for max_target_length, batchdispenser.py:45-49), again for
        # every batch (:81) and again for the prior (:139).  Parsing a 500-label string costs more than reading the
        # utterance's features, so the vectors of the first pass are kept: 4 bytes per frame, read-only.
        self._encoded = {}
        self.max_target_length = max(self._targets(utt_id).size for utt_id in self.target_dict)

    def _targets(self, utt_id):
        """encoded (uint32, read-only) target vector of an utterance, parsed once"""
        enc = self._encoded.get(utt_id)
        if enc is None:
            enc = self.target_coder.encode(self.target_dict[utt_id])
            enc.flags.writeable = False
            self._encoded[utt_id] = enc
        return enc

    def get_batch(self):
        """(list of [T_u, I] float32, list of uint32 [T_u]); utterances without targets or too short to
        splice are skipped with the reference's warnings   (batchdispenser.py:60-91)"""
        batch_inputs, batch_targets = [], []
        while len(batch_inputs) < self.size:
            utt_id, utt_mat, _ = self.feature_reader.get_utt()
            has_targets = utt_id in self.target_dict
            if has_targets and utt_mat is not None:
                batch_inputs.append(utt_mat)
                batch_targets.append(self._targets(utt_id))
                continue
            if not has_targets:
                print("WARNING no targets for %s" % utt_id)
            if utt_mat is None:
                print("WARNING %s is too short to splice" % utt_id)
        return batch_inputs, batch_targets

    def get_raw_batch(self):
        """like get_batch, for the device-side feeder (Trainer.update_raw): returns
        (raw [T_u, D] matrices, their speakers' CMVN statistics, encoded targets); CMVN and splicing are
        left to the GPU.  Same utterance selection and warnings as get_batch."""
        reader = self.feature_reader
        width = 1 + 2 * reader.context_width
        mats, stats, batch_targets = [], [], []
        while len(mats) < self.size:
            utt_id, utt_mat, _ = reader.reader.read_next_utt()
            has_targets = utt_id in self.target_dict
            long_enough = utt_mat.shape[0] >= width
            if has_targets and long_enough:
                mats.append(utt_mat)
                stats.append(reader._cmvn_stats(reader.utt2spk[utt_id]))
                batch_targets.append(self._targets(utt_id))
                continue
            if not has_targets:
                print("WARNING no targets for %s" % utt_id)
            if not long_enough:
                print("WARNING %s is too short to splice" % utt_id)
        return mats, stats, batch_targets

    def split(self):
        """split off what was read so far (validation set)"""
        self.feature_reader.split()

    def _move(self, step):
        moved = 0
        while moved < self.size:
            if step() in self.target_dict:
                moved += 1

    def skip_batch(self):
        self._move(self.feature_reader.next_id)

    def return_batch(self):
        self._move(self.feature_reader.prev_id)

    def compute_target_count(self):
        """occurrences of every label over ALL targets (batchdispenser.py:128-145)"""
        stacked = np.concatenate([self._targets(utt_id) for utt_id in self.target_dict])
        return np.bincount(stacked, minlength=self.target_coder.num_labels)

    @property
    def num_batches(self):
        return self.num_utt // self.size  # Python-2 integer division in the reference

    @property
    def num_utt(self):
        return len(self.target_dict)

    @property
    def num_labels(self):
        return self.target_coder.num_labels

    @property
    def max_input_length(self):
        return self.feature_reader.max_input_length


class AlignmentBatchDispenser(BatchDispenser):
    """targets = gzip'ed `utt pdf1 pdf2 ...` lines from ali-to-pdf (batchdispenser.py:203-223)"""

    def read_target_file(self, target_path):
        table = {}
        with gzip.open(target_path, "rt") as fid:
            for line in fid:
                fields = line.strip().split(" ")
                table[fields[0]] = " ".join(fields[1:])
        return table
