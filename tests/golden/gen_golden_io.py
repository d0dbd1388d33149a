#!/usr/bin/env python
"""Generate tests/golden/io_golden.npz + golden ark/scp bytes BY RUNNING THE REFERENCE'S OWN CODE.

The numpy-only half of the reference (processing/ark.py, feature_reader.py, target_coder.py,
batchdispenser.py, readfiles.py) is Python-2 source.  This script loads those files from
/root/reference, applies the mechanical Python-2 -> Python-3 edits listed in PATCHES (nothing else is
touched; the arithmetic and the byte layout are the reference's), executes them, and records their
outputs on a small deterministic corpus.  Run it in the build container only (it needs
/root/reference); the committed fixtures are what travels to the GPU box.

    python tests/golden/gen_golden_io.py
"""
import gzip
import io
import os
import re
import shutil
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference/processing"
OUT = os.path.dirname(os.path.abspath(__file__))

# (file, old, new) — textual, each one a Python-2-ism
PATCHES = {
    "ark.py": [
        ("np.set_printoptions(threshold=np.nan)\n", ""),  # numpy >= 1.14 rejects nan thresholds
        ("np.set_printoptions(linewidth=np.nan)\n", ""),
        ('print "Input .ark file is not binary"', 'print("Input .ark file is not binary")'),
        ('print "Input .ark file is compressed"', 'print("Input .ark file is compressed")'),
        ('header[0] != "B"', 'header[0] != b"B"'),  # struct 'c' yields bytes on py3
        ('header[1] == "C"', 'header[1] == b"C"'),
        ('header[1] == "F"', 'header[1] == b"F"'),
        ('header[1] == "D"', 'header[1] == b"D"'),
        ("struct.pack('<%ds'%(len(utt_id)), utt_id)", "struct.pack('<%ds'%(len(utt_id)), utt_id.encode())"),
        ("struct.pack('<xcccc', 'B', 'F', 'M', ' ')", "struct.pack('<xcccc', b'B', b'F', b'M', b' ')"),
    ],
    "feature_reader.py": [("import ark\n", "from refproc import ark\n"), ("import readfiles\n", "from refproc import readfiles\n")],
    "batchdispenser.py": [
        ("print 'WARNING no targets for %s' % utt_id", "print('WARNING no targets for %s' % utt_id)"),
        ("print 'WARNING %s is too short to splice' % utt_id", "print('WARNING %s is too short to splice' % utt_id)"),
        ("return self.num_utt/self.size", "return self.num_utt//self.size"),  # py2 int division
        ("gzip.open(target_path, 'rb')", "gzip.open(target_path, 'rt')"),  # str lines as on py2
    ],
    "target_coder.py": [],
    "readfiles.py": [],
}


def load_reference_processing():
    pkg = types.ModuleType("refproc")
    pkg.__path__ = []
    sys.modules["refproc"] = pkg
    mods = {}
    for fname in ["readfiles.py", "ark.py", "target_coder.py", "feature_reader.py", "batchdispenser.py"]:
        src = open(os.path.join(REF, fname)).read()
        for old, new in PATCHES[fname]:
            assert old in src, (fname, old)
            src = src.replace(old, new)
        name = fname[:-3]
        mod = types.ModuleType("refproc." + name)
        sys.modules["refproc." + name] = mod
        setattr(pkg, name, mod)
        exec(compile(src, os.path.join(REF, fname), "exec"), mod.__dict__)
        mods[name] = mod
    return mods


def main():
    ref = load_reference_processing()
    rng = np.random.default_rng(20260925)
    tmp = tempfile.mkdtemp(prefix="tfk_golden_")
    cwd = os.getcwd()
    os.chdir(tmp)  # keep ark paths inside the scp relative and stable
    try:
        D, K = 5, 2
        lengths = {"spkA_utt1": 12, "spkA_utt2": 4, "spkB_utt3": 15, "spkB_utt4": 9, "spkA_utt5": 7, "spkB_utt6": 6}
        feats = {u: rng.standard_normal((t, D)).astype(np.float32) * 2 + 1 for u, t in lengths.items()}
        order = list(lengths)
        # --- ArkWriter (ark.py:190-211): feature archive
        w = ref["ark"].ArkWriter("feats.scp", "feats.ark")
        for u in order:
            w.write_next_utt(u, feats[u])
        w.close()
        # --- per-speaker cmvn statistics in the layout of prepare_data.py:114-117 ([2, D+1])
        stats = {}
        for spk in ("spkA", "spkB"):
            allf = np.concatenate([feats[u] for u in order if u.startswith(spk)])
            s = np.zeros((2, D + 1), np.float32)
            s[0, :-1] = allf.sum(0)
            s[0, -1] = allf.shape[0]
            s[1, :-1] = np.square(allf).sum(0)
            stats[spk] = s
        w = ref["ark"].ArkWriter("cmvn.scp", "cmvn.ark")
        for spk, s in stats.items():
            w.write_next_utt(spk, s)
        w.close()
        with open("utt2spk", "w") as f:
            for u in order:
                f.write("%s %s\n" % (u, u.split("_")[0]))
        # alignments: utt4 has no target on purpose ("WARNING no targets")
        ali = {u: " ".join(str(int(v)) for v in rng.integers(0, 11, lengths[u])) for u in order if u != "spkB_utt4"}
        with gzip.open("pdf.all.gz", "wt") as f:
            for u, a in ali.items():
                f.write("%s %s\n" % (u, a))

        gold = {}
        for name in ("feats.ark", "feats.scp", "cmvn.ark", "cmvn.scp"):
            gold["file_" + name] = np.frombuffer(open(name, "rb").read(), dtype=np.uint8)
        # --- ArkReader (ark.py:28-165)
        r = ref["ark"].ArkReader("feats.scp")
        seq = []
        for _ in range(8):  # wraps around after 6
            uid, mat, looped = r.read_next_utt()
            seq.append("%s:%d:%d" % (uid, mat.shape[0], int(looped)))
        gold["reader_sequence"] = np.array(seq)
        gold["reader_read_utt_spkB_utt3"] = np.array(r.read_utt("spkB_utt3"))
        r2 = ref["ark"].ArkReader("feats.scp")
        r2.read_next_utt(); r2.read_next_utt()
        r2.split()  # drops the read prefix AND the last utterance, cursor not reset (ark.py:161-165)
        gold["split_utt_ids"] = np.array(r2.utt_ids)
        gold["split_scp_position"] = np.array(r2.scp_position)
        ids = [r2.read_next_scp() for _ in range(5)]
        ids += ["prev:" + r2.read_previous_scp() for _ in range(3)]
        gold["split_cursor_walk"] = np.array(ids)
        # --- apply_cmvn / splice (feature_reader.py:91-156)
        fr = ref["feature_reader"]
        for u in order:
            c = fr.apply_cmvn(feats[u], stats[u.split("_")[0]])
            gold["cmvn_" + u] = np.array(c)
            for k in (0, 1, K, 3):
                s = fr.splice(c, k)
                gold["splice%d_%s" % (k, u)] = np.array([]) if s is None else s
                gold["splice%d_none_%s" % (k, u)] = np.array(s is None)
        # --- FeatureReader + AlignmentBatchDispenser (batchdispenser.py)
        reader = fr.FeatureReader("feats.scp", "cmvn.scp", "utt2spk", K, 15)
        coder = ref["target_coder"].AlignmentCoder(lambda x, y: x, 11)
        disp = ref["batchdispenser"].AlignmentBatchDispenser(reader, coder, 2, "pdf.all.gz")
        gold["disp_max_target_length"] = np.array(disp.max_target_length)
        gold["disp_num_batches"] = np.array(disp.num_batches)
        gold["disp_num_utt"] = np.array(disp.num_utt)
        gold["disp_target_count"] = np.array(disp.compute_target_count())
        log = []
        stdout, sys.stdout = sys.stdout, io.StringIO()
        try:
            for b in range(3):
                x, y = disp.get_batch()
                for i, (xi, yi) in enumerate(zip(x, y)):
                    gold["batch%d_x%d" % (b, i)] = xi
                    gold["batch%d_y%d" % (b, i)] = yi
                log.append("batch%d:cursor=%d" % (b, reader.reader.scp_position))
            disp.return_batch()
            log.append("return:cursor=%d" % reader.reader.scp_position)
            disp.skip_batch()
            log.append("skip:cursor=%d" % reader.reader.scp_position)
            warnings = sys.stdout.getvalue()
        finally:
            sys.stdout = stdout
        gold["disp_log"] = np.array(log)
        gold["disp_warnings"] = np.array(warnings)
        enc = coder.encode("3 0 10 10 7")
        gold["encode_example"] = enc
        gold["encode_dtype"] = np.array(str(enc.dtype))
        gold["utt2spk_keys"] = np.array(sorted(ref["readfiles"].read_utt2spk("utt2spk").items()))
        # raw inputs so the tests can rebuild the corpus with the product code
        for u in order:
            gold["feat_" + u] = feats[u]
        for spk, s in stats.items():
            gold["stats_" + spk] = s
        gold["order"] = np.array(order)
        gold["ali_items"] = np.array(sorted(ali.items()))
        np.savez_compressed(os.path.join(OUT, "io_golden.npz"), **gold)
        print("wrote", os.path.join(OUT, "io_golden.npz"), "with", len(gold), "entries")
    finally:
        os.chdir(cwd)
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
