"""TensorFlow checkpoint import/export without TensorFlow (tfkaldi_b200/neuralNetworks/tf_checkpoint.py).

The reference leaves its models as tf.train.Saver checkpoints (neuralNetworks/trainer.py:448-486, decoder.py:73-81) and
ships none, and TensorFlow is not installed here: the container formats are restated from their published layouts and
pinned by (i) known-answer vectors of the primitives (CRC-32C: RFC 3720 B.4; LevelDB's masked-CRC and footer magic;
protobuf varints; a hand-assembled snappy stream), (ii) round trips through both writers, (iii) hand-assembled files
that exercise what the writers never emit (snappy blocks, multi-slice V1 tensors, non-packed repeated fields)."""
import os
import struct

import numpy as np
import pytest

from tfkaldi_b200.neuralNetworks import tf_checkpoint as T


def model(rng, layers=2, bn=True):
    a = {}
    dims = [440] + [64] * layers + [183]
    for l in range(layers + 1):
        a["Classifier/layer%d/parameters/weights" % l] = rng.standard_normal((dims[l], dims[l + 1])).astype(np.float32)
        a["Classifier/layer%d/parameters/biases" % l] = rng.standard_normal(dims[l + 1]).astype(np.float32)
        if bn and l < layers:
            for n in ("beta", "moving_mean", "moving_variance"):
                a["Classifier/layer%d/activation/batch_norm/%s" % (l, n)] = rng.random(dims[l + 1]).astype(np.float32)
    a["Classifier/initialisedlayers"] = np.array(1, np.int32)
    return a


def test_crc32c_known_answers_and_lane_path():
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA  # RFC 3720 B.4
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert T.crc32c(b"") == 0
    rng = np.random.default_rng(0)
    for n in (65536, 65537, 300001, 1 << 20):  # the vectorised lanes + tail against the byte-at-a-time loop
        buf = rng.integers(0, 256, n, dtype=np.uint8)
        assert T.crc32c(buf) == T._crc_bytes(0xFFFFFFFF, buf.tobytes()) ^ 0xFFFFFFFF
    assert T.crc32c(np.arange(70000, dtype=np.float32)) == T.crc32c(np.arange(70000, dtype=np.float32).tobytes())
    # LevelDB crc32c_test: Mask(crc) differs from crc, and masking twice differs again
    c = T.crc32c(b"foo")
    assert T.mask_crc(c) != c and T.mask_crc(T.mask_crc(c)) != c
    assert T.mask_crc(0) == 0xA282EAD8


def test_varints_and_snappy():
    for v in (0, 1, 127, 128, 300, 2 ** 32 - 1, 2 ** 63):
        enc = T._put_varint(v)
        assert T._varint(enc, 0) == (v, len(enc))
    assert T._put_varint(300) == b"\xac\x02"  # protobuf encoding guide
    assert T._signed(T._varint(T._put_varint(-1), 0)[0]) == -1 and len(T._put_varint(-1)) == 10
    # 'abc' literal, then a 1-byte-offset copy of 9 bytes from 3 back (overlapping its own output)
    stream = bytes([12, (3 - 1) << 2]) + b"abc" + bytes([((9 - 4) << 2) | 1, 3])
    assert T.snappy_decompress(stream) == b"abcabcabcabc"
    # 2-byte-offset copy and a long literal (length byte 60 -> one extra length byte)
    lit = bytes(range(100))
    stream = bytes([104, 60 << 2, 99]) + lit + bytes([((4 - 1) << 2) | 2, 100, 0])
    assert T.snappy_decompress(stream) == lit + lit[:4]
    with pytest.raises(ValueError):
        T.snappy_decompress(bytes([5, 0 << 2]) + b"a")  # announces 5 bytes, holds 1


@pytest.mark.parametrize("version", ["v1", "v2"])
def test_round_trip(tmp_path, version):
    rng = np.random.default_rng(1)
    arrays = model(rng)
    arrays["train_variables/global_step"] = np.array(1234, np.int32)
    arrays["train_variables/learning_rate_fact"] = np.array(0.25, np.float32)
    arrays["misc/negatives"] = np.array([-1, 5, -100000], np.int32)
    arrays["misc/doubles"] = rng.standard_normal((3, 2, 2))
    arrays["misc/int64"] = np.array([[1 << 40, -5]], np.int64)
    arrays["misc/empty"] = np.zeros((0, 4), np.float32)
    for i in range(300):  # enough keys for several index entries / prefix-compressed blocks
        arrays["many/v%03d" % i] = np.full(3, i, np.float32)
    path = str(tmp_path / "final")
    (T.write_v1 if version == "v1" else T.write_v2)(path, arrays)
    assert T.find(path) == version
    got = T.read(path)
    assert sorted(got) == sorted(arrays)
    for k, want in arrays.items():
        assert got[k].dtype == want.dtype and got[k].shape == want.shape and np.array_equal(got[k], want), k
    if version == "v2":
        assert os.path.getsize(path + ".data-00000-of-00001") == sum(a.nbytes for a in arrays.values())
    # the table ends with LevelDB's magic number
    with open(path if version == "v1" else path + ".index", "rb") as f:
        assert f.read()[-8:] == struct.pack("<Q", 0xDB4775248B80FB57)


def test_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(2)
    arrays = model(rng, bn=False)
    path = str(tmp_path / "m")
    T.write_v2(path, arrays)
    data = path + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read())
    raw[1000] ^= 0x10
    open(data, "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        T.read(path)
    assert set(T.read(path, verify=False)) == set(arrays)  # readable when told not to check
    T.write_v1(path + "1", arrays)
    raw = bytearray(open(path + "1", "rb").read())
    raw[len(raw) // 2] ^= 0x01
    open(path + "1", "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        T.read(path + "1")
    open(path + "2", "wb").write(b"not a checkpoint at all, but longer than forty-eight bytes ........")
    assert T.find(path + "2") is None
    with pytest.raises(FileNotFoundError):
        T.read(path + "2")


def _block(entries):
    body = b"".join(T._put_varint(0) + T._put_varint(len(k)) + T._put_varint(len(v)) + k + v for k, v in entries)
    restarts = b"".join(struct.pack("<I", 0) for _ in range(1))
    return body + restarts + struct.pack("<I", 1)


def _snappy_literal(raw):
    """a valid snappy stream made of literals only"""
    out = T._put_varint(len(raw))
    for i in range(0, len(raw), 60):
        chunk = raw[i:i + 60]
        out += bytes([(len(chunk) - 1) << 2]) + chunk
    return out


def test_hand_assembled_v1_with_snappy_block_slices_and_unpacked_fields(tmp_path):
    """What TensorSliceWriter can emit but write_v1 never does: a snappy-compressed data block, one tensor stored as
    two row slices (partitioned save), float_val as non-packed repeated fields, and a scalar int_val."""
    ld, tag, vi = T._ld, T._tag, T._put_varint
    shape = T._shape_msg((4, 3))
    whole = ld(1, b"") + ld(1, b"")
    meta = ld(1, ld(1, b"w") + ld(2, shape) + tag(3, 0) + vi(1) + ld(4, whole))
    meta += ld(1, ld(1, b"step") + ld(2, b"") + tag(3, 0) + vi(3))
    w = np.arange(12, dtype=np.float32).reshape(4, 3) - 5

    def rows(lo, n):  # TensorSliceProto: rows [lo, lo+n), all columns
        ext = (tag(1, 0) + vi(lo) if lo else b"") + tag(2, 0) + vi(n)
        return ld(1, ext) + ld(1, b"")

    def tensor(vals, packed):
        if packed:
            return tag(1, 0) + vi(1) + ld(5, vals.astype("<f4").tobytes())
        return tag(1, 0) + vi(1) + b"".join(tag(5, 5) + struct.pack("<f", v) for v in vals)

    entries = [
        (b"", ld(1, meta + ld(2, tag(1, 0) + vi(1)))),
        (b"\x00step", ld(2, ld(1, b"step") + ld(3, tag(1, 0) + vi(3) + tag(7, 0) + vi(77)))),
        (b"\x00w-a", ld(2, ld(1, b"w") + ld(2, rows(0, 1)) + ld(3, tensor(w[:1].ravel(), False)))),
        (b"\x00w-b", ld(2, ld(1, b"w") + ld(2, rows(1, 3)) + ld(3, tensor(w[1:].ravel(), True)))),
    ]
    path = str(tmp_path / "ckpt")
    with open(path, "wb") as f:
        def emit(body, ctype):
            raw = body + bytes([ctype])
            off = f.tell()
            f.write(raw + struct.pack("<I", T.mask_crc(T.crc32c(raw))))
            return vi(off) + vi(len(body))

        h_data = emit(_snappy_literal(_block(entries)), 1)
        h_meta = emit(_block([]), 0)
        h_index = emit(_block([(entries[-1][0], h_data)]), 0)
        footer = h_meta + h_index
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", T.MAGIC))
    got = T.read(path)
    assert np.array_equal(got["w"], w) and got["w"].dtype == np.float32
    assert got["step"].shape == () and int(got["step"]) == 77 and got["step"].dtype == np.int32


def test_model_file_lookup_prefers_npz_and_falls_back_to_tensorflow(tmp_path):
    """Trainer.restore_model / Decoder.restore read `<filename>.npz`, else the reference's checkpoint at `filename`."""
    from tfkaldi_b200.neuralNetworks.trainer import MODEL_NAMES, read_model_file

    rng = np.random.default_rng(3)
    arrays = model(rng)
    ref_trained = str(tmp_path / "final")
    T.write_v2(ref_trained, arrays)
    got = read_model_file(ref_trained)
    assert all(np.array_equal(got[k], v) for k, v in arrays.items())
    np.savez(ref_trained + ".npz", **{"Classifier/layer0/parameters/biases": np.ones(3, np.float32)})
    assert list(read_model_file(ref_trained)) == ["Classifier/layer0/parameters/biases"]
    with pytest.raises(FileNotFoundError):
        read_model_file(str(tmp_path / "missing"))
    # every classifier variable the reference creates maps onto an engine tensor
    for key in arrays:
        if key != "Classifier/initialisedlayers":
            assert key.split("/", 2)[2] in MODEL_NAMES, key
