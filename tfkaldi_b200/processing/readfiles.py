"""Kaldi text-table readers used on the hot path (reference: processing/readfiles.py:89-105)."""


def read_utt2spk(filename):
    """utterance id -> speaker id, from lines `utt spk`."""
    table = {}
    with open(filename) as fid:
        for line in fid:
            fields = line.replace("\n", "").split(" ")
            table[fields[0]] = fields[1]
    return table
