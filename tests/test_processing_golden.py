"""The product's processing layer against golden vectors produced by the REFERENCE'S OWN code
(tests/golden/gen_golden_io.py ran processing/*.py from /root/reference): byte-exact archives,
bit-exact CMVN / splice / encoding, identical cursor behaviour including the reference's quirks."""
import gzip
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "io_golden.npz"), allow_pickle=False)


@pytest.fixture()
def corpus(tmp_path, monkeypatch):
    from tfkaldi_b200.processing import ark

    monkeypatch.chdir(tmp_path)
    order = [str(u) for u in GOLD["order"]]
    w = ark.ArkWriter("feats.scp", "feats.ark")
    for u in order:
        w.write_next_utt(u, GOLD["feat_" + u])
    w.close()
    w = ark.ArkWriter("cmvn.scp", "cmvn.ark")
    for spk in ("spkA", "spkB"):
        w.write_next_utt(spk, GOLD["stats_" + spk])
    w.close()
    with open("utt2spk", "w") as f:
        for u in order:
            f.write("%s %s\n" % (u, u.split("_")[0]))
    with gzip.open("pdf.all.gz", "wt") as f:
        for u, a in GOLD["ali_items"]:
            f.write("%s %s\n" % (u, a))
    return order


def test_arkwriter_bytes_identical_to_reference(corpus):
    for name in ("feats.ark", "feats.scp", "cmvn.ark", "cmvn.scp"):
        assert open(name, "rb").read() == bytes(GOLD["file_" + name]), name


def test_arkwriter_appends_like_the_reference(corpus):
    from tfkaldi_b200.processing import ark

    w = ark.ArkWriter("more.scp", "feats.ark")  # append to an existing archive (ark.py:201)
    w.write_next_utt("late", np.ones((2, 5)))
    w.close()
    size = len(bytes(GOLD["file_feats.ark"]))
    assert open("more.scp").read() == "late feats.ark:%d\n" % (size + 4)
    r = ark.ArkReader("more.scp")
    assert np.array_equal(r.read_utt("late"), np.ones((2, 5), np.float32))


def test_arkreader_matches_reference(corpus):
    from tfkaldi_b200.processing import ark

    r = ark.ArkReader("feats.scp")
    seq = []
    for _ in range(8):
        uid, mat, looped = r.read_next_utt()
        assert np.array_equal(mat, GOLD["feat_" + uid]) and mat.dtype == np.float32
        seq.append("%s:%d:%d" % (uid, mat.shape[0], int(looped)))
    assert seq == [str(s) for s in GOLD["reader_sequence"]]
    assert np.array_equal(r.read_utt("spkB_utt3"), GOLD["reader_read_utt_spkB_utt3"])
    r2 = ark.ArkReader("feats.scp")
    r2.read_next_utt()
    r2.read_next_utt()
    r2.split()
    assert r2.utt_ids == [str(s) for s in GOLD["split_utt_ids"]] and r2.scp_position == int(GOLD["split_scp_position"])
    walk = [r2.read_next_scp() for _ in range(5)] + ["prev:" + r2.read_previous_scp() for _ in range(3)]
    assert walk == [str(s) for s in GOLD["split_cursor_walk"]]


def test_arkreader_reads_double_archives_and_rejects_others(tmp_path):
    import struct

    from tfkaldi_b200.processing import ark

    m = np.arange(6, dtype=np.float64).reshape(2, 3)
    with open(tmp_path / "d.ark", "wb") as f:
        f.write(b"k" + struct.pack("<xcccc", b"B", b"D", b"M", b" ") + struct.pack("<bi", 4, 2) + struct.pack("<bi", 4, 3) + m.tobytes())
        bad = f.tell()
        f.write(b"t [ 1 2 ]  padding so the header reads stay inside the file")
        comp = f.tell() + 1
        f.write(b"c" + struct.pack("<xcccc", b"B", b"C", b"M", b" ") + b"\0" * 16)
    with open(tmp_path / "d.scp", "w") as f:
        f.write("k %s:1\nt %s:%d\nc %s:%d\n" % (tmp_path / "d.ark", tmp_path / "d.ark", bad, tmp_path / "d.ark", comp))
    r = ark.ArkReader(str(tmp_path / "d.scp"))
    assert np.array_equal(r.read_utt("k"), m) and r.read_utt("k").dtype == np.float64
    with pytest.raises(SystemExit):
        r.read_utt("t")  # "Input .ark file is not binary" + exit(1)   ark.py:73-75
    with pytest.raises(SystemExit):
        r.read_utt("c")  # compressed   ark.py:76-78


def test_cmvn_and_splice_bit_exact(corpus):
    from tfkaldi_b200.processing import feature_reader as fr

    for u in corpus:
        c = fr.apply_cmvn(GOLD["feat_" + u], GOLD["stats_" + u.split("_")[0]])
        assert c.dtype == GOLD["cmvn_" + u].dtype and np.array_equal(c, GOLD["cmvn_" + u])
        for k in (0, 1, 2, 3):
            s = fr.splice(c, k)
            if bool(GOLD["splice%d_none_%s" % (k, u)]):
                assert s is None
            else:
                g = GOLD["splice%d_%s" % (k, u)]
                assert s.dtype == np.float32 and s.shape == g.shape and np.array_equal(s, g)


def test_batchdispenser_matches_reference(corpus, capsys):
    from tfkaldi_b200.processing import batchdispenser, feature_reader, target_coder

    reader = feature_reader.FeatureReader("feats.scp", "cmvn.scp", "utt2spk", 2, 15)
    coder = target_coder.AlignmentCoder(lambda x, y: x, 11)
    disp = batchdispenser.AlignmentBatchDispenser(reader, coder, 2, "pdf.all.gz")
    assert disp.max_target_length == int(GOLD["disp_max_target_length"])
    assert disp.num_batches == int(GOLD["disp_num_batches"]) and disp.num_utt == int(GOLD["disp_num_utt"])
    assert disp.num_labels == 11 and disp.max_input_length == 15
    assert np.array_equal(disp.compute_target_count(), GOLD["disp_target_count"])
    log = []
    capsys.readouterr()
    for b in range(3):
        x, y = disp.get_batch()
        assert len(x) == 2
        for i in range(2):
            assert np.array_equal(x[i], GOLD["batch%d_x%d" % (b, i)]) and x[i].dtype == np.float32
            assert np.array_equal(y[i], GOLD["batch%d_y%d" % (b, i)]) and y[i].dtype == np.uint32
        log.append("batch%d:cursor=%d" % (b, reader.reader.scp_position))
    disp.return_batch()
    log.append("return:cursor=%d" % reader.reader.scp_position)
    disp.skip_batch()
    log.append("skip:cursor=%d" % reader.reader.scp_position)
    assert log == [str(s) for s in GOLD["disp_log"]]
    assert capsys.readouterr().out == str(GOLD["disp_warnings"])


def test_alignment_coder(corpus):
    from tfkaldi_b200.processing import readfiles, target_coder

    coder = target_coder.AlignmentCoder(lambda x, y: x, 11)
    enc = coder.encode("3 0 10 10 7")
    assert str(enc.dtype) == str(GOLD["encode_dtype"]) and np.array_equal(enc, GOLD["encode_example"])
    assert coder.decode(enc) == "3 0 10 10 7"
    with pytest.raises(KeyError):
        coder.encode("3 11")  # not in the alphabet
    assert sorted(readfiles.read_utt2spk("utt2spk").items()) == [tuple(map(str, kv)) for kv in GOLD["utt2spk_keys"]]


def test_raw_batches_select_the_same_utterances(corpus, capsys):
    """get_raw_batch (device-side feeder) walks the archive exactly like get_batch: same utterances, same
    targets, same warnings; cmvn+splice of its raw matrices reproduces get_batch's features bit for bit."""
    from tfkaldi_b200.processing import batchdispenser, feature_reader, target_coder

    def make():
        reader = feature_reader.FeatureReader("feats.scp", "cmvn.scp", "utt2spk", 2, 15)
        return reader, batchdispenser.AlignmentBatchDispenser(reader, target_coder.AlignmentCoder(lambda x, y: x, 11), 2, "pdf.all.gz")

    _, a = make()
    _, b = make()
    capsys.readouterr()
    for _ in range(3):
        x, y = a.get_batch()
        out_a = capsys.readouterr().out
        raw, stats, y2 = b.get_raw_batch()
        assert capsys.readouterr().out == out_a
        for xi, ri, si, ya, yb in zip(x, raw, stats, y, y2):
            assert np.array_equal(feature_reader.splice(feature_reader.apply_cmvn(ri, si), 2), xi)
            assert np.array_equal(ya, yb)
