"""Feature reading with per-speaker CMVN and +-k frame splicing
(reference: processing/feature_reader.py).  Same API and values; the splice is a single strided
window over a zero-padded copy instead of 2k slice assignments, and CMVN statistics are looked up
once per speaker instead of being re-read for every utterance."""
import numpy as np

from . import ark, readfiles


class FeatureReader(object):
    def __init__(self, scpfile, cmvnfile, utt2spkfile, context_width, max_input_length):
        self.reader = ark.ArkReader(scpfile)
        self.reader_cmvn = ark.ArkReader(cmvnfile)
        self.utt2spk = readfiles.read_utt2spk(utt2spkfile)
        self.context_width = context_width
        self.max_input_length = max_input_length
        self._stats = {}

    def _cmvn_stats(self, speaker):
        stats = self._stats.get(speaker)
        if stats is None:
            stats = np.array(self.reader_cmvn.read_utt(speaker))
            self._stats[speaker] = stats
        return stats

    def get_utt(self):
        """(utt_id, normalised+spliced features or None if too short, looped)   feature_reader.py:42-60"""
        utt_id, utt_mat, looped = self.reader.read_next_utt()
        utt_mat = apply_cmvn(utt_mat, self._cmvn_stats(self.utt2spk[utt_id]))
        return utt_id, splice(utt_mat, self.context_width), looped

    def get_utt_raw(self):
        """(utt_id, CMVN-normalised UNSPLICED features, looped): for the device-side splicer."""
        utt_id, utt_mat, looped = self.reader.read_next_utt()
        return utt_id, apply_cmvn(utt_mat, self._cmvn_stats(self.utt2spk[utt_id])), looped

    def get_raw_utt(self):
        """(utt_id, UN-normalised unspliced features, the speaker's CMVN statistics, looped): for the decoder's
        device-side CMVN + splice (tfk_forward_loglik_raw)"""
        utt_id, utt_mat, looped = self.reader.read_next_utt()
        return utt_id, utt_mat, (None if utt_id is None else self._cmvn_stats(self.utt2spk[utt_id])), looped

    def next_id(self):
        return self.reader.read_next_scp()

    def prev_id(self):
        return self.reader.read_previous_scp()

    def split(self):
        self.reader.split()


def apply_cmvn(utt, stats):
    """mean/variance normalisation from accumulated statistics (feature_reader.py:91-115):
    stats[0] = [sum(x), count], stats[1] = [sum(x^2), 0].  No variance floor (reference behaviour)."""
    count = stats[0, -1]
    mean = stats[0, :-1] / count
    variance = stats[1, :-1] / count - np.square(mean)
    return np.divide(np.subtract(utt, mean), np.sqrt(variance))


def splice(utt, context_width):
    """[T, D] -> float32 [T, D*(2k+1)], row t = [x[t-k] .. x[t] .. x[t+k]] with zeros beyond the
    utterance edges; None when T < 2k+1   (feature_reader.py:117-156)."""
    frames, dim = utt.shape
    width = 1 + 2 * context_width
    if frames < width:
        return None
    padded = np.zeros((frames + 2 * context_width, dim), dtype=np.float32)
    padded[context_width:context_width + frames] = utt
    windows = np.lib.stride_tricks.sliding_window_view(padded, (width, dim))[:, 0]
    return np.ascontiguousarray(windows).reshape(frames, width * dim)
