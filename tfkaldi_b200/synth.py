"""Synthetic Kaldi-style corpora in the exact file layouts the reference consumes (there is no
network for real data): feats.ark/scp + feats_shuffled.scp + maxlength (processing/prepare_data.py:13-80,
124-141), per-speaker CMVN statistics [2, D+1] (prepare_data.py:82-122), utt2spk / spk2utt, and
gzip'ed `utt pdf pdf ...` alignments as written by `ali-to-pdf | gzip` (kaldi/gmm.py:71-76)."""
import gzip
import os

import numpy as np

from .processing import ark


def make_corpus(featdir, num_utts=64, min_len=200, max_len=400, feat_dim=40, num_speakers=4, num_pdfs=183,
                seed=0, alidir=None, shuffle=True):
    """write a corpus under `featdir`; returns a dict describing it (paths, lengths, num_pdfs)"""
    rng = np.random.default_rng(seed)
    os.makedirs(featdir, exist_ok=True)
    alidir = alidir or featdir
    os.makedirs(alidir, exist_ok=True)
    cwd = os.getcwd()
    for stale in ("feats.ark", "cmvn.ark"):
        path = os.path.join(featdir, stale)
        if os.path.exists(path):
            os.remove(path)  # ArkWriter appends (ark.py:201); the caller removes stale archives (main.py:177-178)
    utts, lengths, spk_of = [], {}, {}
    feats = {}
    for i in range(num_utts):
        spk = "spk%02d" % (i % num_speakers)
        utt = "%s_utt%04d" % (spk, i)
        t = int(rng.integers(min_len, max_len + 1))
        # speaker-dependent offset / scale so CMVN has something to do
        x = rng.standard_normal((t, feat_dim)).astype(np.float32) * (1.0 + 0.1 * (i % num_speakers)) + 0.5 * (i % num_speakers)
        utts.append(utt); lengths[utt] = t; spk_of[utt] = spk; feats[utt] = x
    writer = ark.ArkWriter(os.path.join(featdir, "feats.scp"), os.path.join(featdir, "feats.ark"))
    for utt in utts:
        writer.write_next_utt(utt, feats[utt])
    writer.close()
    spk2utt = {}
    for utt in utts:
        spk2utt.setdefault(spk_of[utt], []).append(utt)
    with open(os.path.join(featdir, "utt2spk"), "w") as f:
        for utt in utts:
            f.write("%s %s\n" % (utt, spk_of[utt]))
    with open(os.path.join(featdir, "spk2utt"), "w") as f:
        for spk, us in spk2utt.items():
            f.write("%s %s\n" % (spk, " ".join(us)))
    writer = ark.ArkWriter(os.path.join(featdir, "cmvn.scp"), os.path.join(featdir, "cmvn.ark"))
    for spk, us in spk2utt.items():
        data = np.concatenate([feats[u] for u in us], axis=0)
        stats = np.zeros([2, feat_dim + 1])
        stats[0, :feat_dim] = np.sum(data, 0)
        stats[1, :feat_dim] = np.sum(np.square(data), 0)
        stats[0, feat_dim] = data.shape[0]
        writer.write_next_utt(spk, stats)
    writer.close()
    with open(os.path.join(featdir, "maxlength"), "w") as f:
        f.write(str(max(lengths.values())))
    lines = open(os.path.join(featdir, "feats.scp")).readlines()
    if shuffle:
        rng.shuffle(lines)
    with open(os.path.join(featdir, "feats_shuffled.scp"), "w") as f:
        f.writelines(lines)
    ali_path = os.path.join(alidir, "pdf.all")
    with gzip.open(ali_path, "wt") as f:
        for utt in utts:
            f.write("%s %s\n" % (utt, " ".join(str(int(v)) for v in rng.integers(0, num_pdfs, lengths[utt]))))
    os.chdir(cwd)
    return {"featdir": featdir, "alifile": ali_path, "utts": utts, "lengths": lengths, "max_length": max(lengths.values()),
            "feat_dim": feat_dim, "num_pdfs": num_pdfs}
