"""TensorFlow checkpoint reader / writer without TensorFlow.

The reference saves its models with `tf.train.Saver().save(session, filename)` (neuralNetworks/trainer.py:448-475,
decoder.py:73-81).  TensorFlow r0.11 writes the V1 "tensor slice" format (one SSTable file), r0.12 defaults to the V2
"tensor bundle" format (`<prefix>.index` SSTable + `<prefix>.data-00000-of-0000N`).  Both are restated here from the
published formats so that models trained by the reference load into this engine (Trainer.restore_model /
Decoder.restore fall back to this module when no .npz is found) and models trained here can be handed back
(`write_v2`).  TensorFlow is not installed in this image and the reference ships no checkpoint, so the format
knowledge is PINNED ONLY BY ROUND TRIPS through this module's own writers (tests/test_tf_checkpoint.py) plus the
known-answer vectors of the primitives (CRC32C, varints, snappy).

Formats (little-endian throughout):
  * SSTable = LevelDB table: blocks of prefix-compressed (shared, non_shared, value_len, key_delta, value) entries
    followed by a restart array, each block trailed by 1 compression byte (0 raw, 1 snappy) + masked CRC32C; an index
    block maps last-keys to block handles; 48-byte footer = metaindex handle, index handle, padding, magic
    0xdb4775248b80fb57.
  * V1: key "" -> SavedTensorSlices{meta}, other keys -> SavedTensorSlices{data: SavedSlice{name, slice, TensorProto}}.
  * V2: key "" -> BundleHeaderProto, key <tensor name> -> BundleEntryProto{dtype, shape, shard_id, offset, size,
    crc32c}; tensor bytes live in the data shard at [offset, offset + size).
"""
from __future__ import annotations

import os
import struct

import numpy as np

MAGIC = 0xDB4775248B80FB57
DT = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8"), 10: np.dtype("bool")}
DT_CODE = {np.dtype("float32"): 1, np.dtype("float64"): 2, np.dtype("int32"): 3, np.dtype("int64"): 9}


# ------------------------------------------------------------------------------------------ primitives
def _crc_table():
    t = np.arange(256, dtype=np.uint32)
    for _ in range(8):
        t = np.where(t & 1, (t >> 1) ^ np.uint32(0x82F63B78), t >> 1).astype(np.uint32)
    return t


_T = _crc_table()
_TL = [int(v) for v in _T]


def _crc_bytes(state: int, data: bytes) -> int:
    for b in data:
        state = _TL[(state ^ b) & 0xFF] ^ (state >> 8)
    return state


def crc32c(data) -> int:
    """CRC-32C (Castagnoli).  Large buffers are cut into equal lanes whose registers advance together in numpy; the
    lanes are then chained with the linear 'advance through n zero bytes' operator (the register update is affine in
    the state), so 100 MB of weights take a fraction of a second instead of minutes of per-byte Python."""
    buf = np.frombuffer(memoryview(data).cast("B"), dtype=np.uint8) if not isinstance(data, np.ndarray) else data.reshape(-1).view(np.uint8)
    n = buf.size
    if n < (1 << 16):
        return _crc_bytes(0xFFFFFFFF, buf.tobytes()) ^ 0xFFFFFFFF
    lanes = 1 << max(4, min(14, int(round(np.log2(np.sqrt(n) / 2)))))
    m = n // lanes
    body = buf[: lanes * m].reshape(lanes, m)
    r = np.zeros(lanes, dtype=np.uint32)  # raw registers from state 0
    basis = (np.uint32(1) << np.arange(32, dtype=np.uint32)).astype(np.uint32)  # to derive the zero-advance operator
    for j in range(m):
        r = _T[(r ^ body[:, j]) & np.uint32(0xFF)] ^ (r >> np.uint32(8))
        basis = _T[basis & np.uint32(0xFF)] ^ (basis >> np.uint32(8))
    cols = [int(v) for v in basis]  # image of bit i after m zero bytes

    def advance(s):
        out, i = 0, 0
        while s:
            if s & 1:
                out ^= cols[i]
            s >>= 1
            i += 1
        return out

    state = 0xFFFFFFFF
    for v in r.tolist():
        state = advance(state) ^ v
    state = _crc_bytes(state, buf[lanes * m:].tobytes())
    return state ^ 0xFFFFFFFF


def mask_crc(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


def _varint(buf, pos):
    out = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7


def _put_varint(v: int) -> bytes:
    v &= (1 << 64) - 1
    out = bytearray()
    while v >= 0x80:
        out.append((v & 0x7F) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def snappy_decompress(data: bytes) -> bytes:
    n, pos = _varint(data, 0)
    out = bytearray()
    while pos < len(data):
        tag = data[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(data[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += data[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln, off = ((tag >> 2) & 7) + 4, ((tag >> 5) << 8) | data[pos]
            pos += 1
        elif kind == 2:
            ln, off = (tag >> 2) + 1, int.from_bytes(data[pos:pos + 2], "little")
            pos += 2
        else:
            ln, off = (tag >> 2) + 1, int.from_bytes(data[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise ValueError("corrupt snappy stream")
        for _ in range(ln):  # copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


# ------------------------------------------------------------------------------------------ protobuf wire format
def _fields(buf):
    """yield (field number, wire type, value) of one message; length-delimited values come back as bytes"""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _varint(buf, pos)
        elif wt == 1:
            val, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            val, pos = bytes(buf[pos:pos + ln]), pos + ln
        elif wt == 5:
            val, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield num, wt, val


def _signed(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _tag(num, wt):
    return _put_varint((num << 3) | wt)


def _ld(num, payload: bytes) -> bytes:
    return _tag(num, 2) + _put_varint(len(payload)) + payload


def _shape_msg(shape) -> bytes:  # TensorShapeProto{ repeated Dim dim = 2 { int64 size = 1 } }
    return b"".join(_ld(2, _tag(1, 0) + _put_varint(int(d))) for d in shape)


def _parse_shape(buf):
    dims = []
    for num, _, val in _fields(buf):
        if num == 2:
            size = 0
            for n2, _, v2 in _fields(val):
                if n2 == 1:
                    size = _signed(v2)
            dims.append(size)
    return tuple(dims)


def _parse_slice(buf, shape):
    """TensorSliceProto{ repeated Extent extent = 1 { int64 start = 1; int64 length = 2 } } -> tuple of python slices"""
    out = []
    for num, _, val in _fields(buf):
        if num == 1:
            start, length = 0, None
            for n2, _, v2 in _fields(val):
                if n2 == 1:
                    start = _signed(v2)
                elif n2 == 2:
                    length = _signed(v2)
            d = len(out)
            out.append(slice(start, shape[d] if length is None or length < 0 else start + length))
    return tuple(out) if out else tuple(slice(0, d) for d in shape)


# ------------------------------------------------------------------------------------------ SSTable
def _read_block(f, offset, size, verify):
    f.seek(offset)
    raw = f.read(size + 5)
    if len(raw) != size + 5:
        raise ValueError("truncated table block")
    body, ctype, crc = raw[:size], raw[size], struct.unpack("<I", raw[size + 1:])[0]
    if verify and mask_crc(crc32c(raw[:size + 1])) != crc:
        raise ValueError("table block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        body = snappy_decompress(body)
    elif ctype != 0:
        raise ValueError("unknown block compression %d" % ctype)
    return body


def _block_entries(body):
    nrestarts = struct.unpack("<I", body[-4:])[0]
    end = len(body) - 4 - 4 * nrestarts
    pos, key = 0, b""
    while pos < end:
        shared, pos = _varint(body, pos)
        non_shared, pos = _varint(body, pos)
        vlen, pos = _varint(body, pos)
        key = key[:shared] + body[pos:pos + non_shared]
        pos += non_shared
        yield key, body[pos:pos + vlen]
        pos += vlen


def read_table(path, verify=True):
    """all (key, value) pairs of a LevelDB-format table file, in key order"""
    out = []
    with open(path, "rb") as f:
        f.seek(0, os.SEEK_END)
        size = f.tell()
        if size < 48:
            raise ValueError("%s: too short for a table file" % path)
        f.seek(size - 48)
        footer = f.read(48)
        if struct.unpack("<Q", footer[40:])[0] != MAGIC:
            raise ValueError("%s: not a TensorFlow/LevelDB table (bad magic)" % path)
        _, pos = _varint(footer, 0)  # metaindex handle (unused)
        _, pos = _varint(footer, pos)
        ioff, pos = _varint(footer, pos)
        isize, pos = _varint(footer, pos)
        for _, handle in _block_entries(_read_block(f, ioff, isize, verify)):
            boff, p = _varint(handle, 0)
            bsize, _ = _varint(handle, p)
            out.extend(_block_entries(_read_block(f, boff, bsize, verify)))
    return out


def write_table(path, items, block_size=4096, restart_interval=16):
    """items: iterable of (key bytes, value bytes); written in sorted key order, uncompressed blocks"""
    items = sorted(items)
    with open(path, "wb") as f:
        offset = 0

        def emit(body):
            nonlocal offset
            raw = body + b"\x00"
            f.write(raw + struct.pack("<I", mask_crc(crc32c(raw))))
            handle = _put_varint(offset) + _put_varint(len(body))
            offset += len(raw) + 4
            return handle

        def build(entries):
            body, restarts, prev = bytearray(), [], b""
            for i, (k, v) in enumerate(entries):
                shared = 0
                if i % restart_interval == 0:
                    restarts.append(len(body))
                else:
                    while shared < min(len(prev), len(k)) and prev[shared] == k[shared]:
                        shared += 1
                body += _put_varint(shared) + _put_varint(len(k) - shared) + _put_varint(len(v)) + k[shared:] + v
                prev = k
            restarts = restarts or [0]
            return bytes(body) + b"".join(struct.pack("<I", r) for r in restarts) + struct.pack("<I", len(restarts))

        index, cur, cur_bytes = [], [], 0
        for k, v in items:
            cur.append((k, v))
            cur_bytes += len(k) + len(v) + 3
            if cur_bytes >= block_size:
                index.append((cur[-1][0], emit(build(cur))))
                cur, cur_bytes = [], 0
        if cur or not index:
            index.append((cur[-1][0] if cur else b"", emit(build(cur))))
        meta = emit(build([]))
        idx = emit(build(index))
        footer = meta + idx
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))


# ------------------------------------------------------------------------------------------ V2: tensor bundle
def read_v2(prefix, verify=True):
    entries = read_table(prefix + ".index", verify)
    num_shards, out, shards = 1, {}, {}
    for key, val in entries:
        if key == b"":
            for num, _, v in _fields(val):  # BundleHeaderProto{ num_shards = 1; endianness = 2; version = 3 }
                if num == 1:
                    num_shards = v
                elif num == 2 and v != 0:
                    raise ValueError("big-endian tensor bundles are not supported")
            continue
        dtype = shape = None
        shard = offset = size = 0
        crc, sliced = None, False
        for num, wt, v in _fields(val):  # BundleEntryProto
            if num == 1:
                dtype = v
            elif num == 2:
                shape = _parse_shape(v)
            elif num == 3:
                shard = v
            elif num == 4:
                offset = v
            elif num == 5:
                size = v
            elif num == 6:
                crc = struct.unpack("<I", v)[0]
            elif num == 7:
                sliced = True
        if sliced:
            raise ValueError("partitioned variable %r: sliced bundle entries are not supported" % key.decode())
        if dtype not in DT:
            continue  # strings etc.: nothing the reference saves
        if shard not in shards:
            shards[shard] = np.memmap("%s.data-%05d-of-%05d" % (prefix, shard, num_shards), dtype=np.uint8, mode="r")
        raw = shards[shard][offset:offset + size]
        if verify and crc is not None and mask_crc(crc32c(np.asarray(raw))) != crc:
            raise ValueError("tensor %r: checksum mismatch" % key.decode())
        out[key.decode()] = np.frombuffer(bytes(raw), dtype=DT[dtype]).reshape(shape or ()).copy()
    return out


def write_v2(prefix, arrays):
    """{name: ndarray} -> <prefix>.index + <prefix>.data-00000-of-00001 (what tf.train.Saver(write_version=V2) emits)"""
    header = _tag(1, 0) + _put_varint(1) + _ld(3, _tag(1, 0) + _put_varint(1))  # num_shards 1, little endian, producer 1
    items, offset = [(b"", header)], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for name in sorted(arrays):
            a = np.asarray(arrays[name], order="C")  # (ascontiguousarray would turn scalars into 1-vectors)
            if a.dtype not in DT_CODE:
                raise ValueError("%s: unsupported dtype %s" % (name, a.dtype))
            raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
            f.write(raw)
            entry = _tag(1, 0) + _put_varint(DT_CODE[a.dtype])
            entry += _ld(2, _shape_msg(a.shape))
            if offset:
                entry += _tag(4, 0) + _put_varint(offset)
            entry += _tag(5, 0) + _put_varint(len(raw)) + _tag(6, 5) + struct.pack("<I", mask_crc(crc32c(raw)))
            items.append((name.encode(), entry))
            offset += len(raw)
    write_table(prefix + ".index", items)


# ------------------------------------------------------------------------------------------ V1: tensor slices
_TP_FIELDS = {5: ("<f4", 5), 6: ("<f8", 1), 7: None, 10: None}  # packed repeated numeric fields of TensorProto


def _parse_tensor_proto(buf, dtype):
    content, vals = None, []
    for num, wt, v in _fields(buf):
        if num == 4:
            content = v  # tensor_content
        elif num == 5:  # float_val
            vals.append(np.frombuffer(v, "<f4") if wt == 2 else np.frombuffer(v, "<f4", count=1))
        elif num == 6:  # double_val
            vals.append(np.frombuffer(v, "<f8") if wt == 2 else np.frombuffer(v, "<f8", count=1))
        elif num in (7, 10, 11):  # int_val / int64_val / bool_val: varints, packed or not
            if wt == 2:
                pos, seq = 0, []
                while pos < len(v):
                    x, pos = _varint(v, pos)
                    seq.append(_signed(x))
                vals.append(np.array(seq, dtype=np.int64))
            else:
                vals.append(np.array([_signed(v)], dtype=np.int64))
    if content is not None:
        return np.frombuffer(content, dtype=dtype)
    return np.concatenate(vals).astype(dtype) if vals else np.zeros(0, dtype)


def read_v1(path, verify=True):
    entries = read_table(path, verify)
    meta = {}
    out = {}
    for key, val in entries:
        for num, _, v in _fields(val):  # SavedTensorSlices{ meta = 1; data = 2 }
            if num == 1 and key == b"":
                for n2, _, v2 in _fields(v):  # SavedTensorSliceMeta{ repeated SavedSliceMeta tensor = 1 }
                    if n2 != 1:
                        continue
                    name, shape, dtype = None, (), None
                    for n3, _, v3 in _fields(v2):  # SavedSliceMeta{ name = 1; shape = 2; type = 3; slice = 4 }
                        if n3 == 1:
                            name = v3.decode()
                        elif n3 == 2:
                            shape = _parse_shape(v3)
                        elif n3 == 3:
                            dtype = v3
                    meta[name] = (shape, dtype)
            elif num == 2:
                name, sl, tp = None, b"", b""
                for n2, _, v2 in _fields(v):  # SavedSlice{ name = 1; slice = 2; data = 3 }
                    if n2 == 1:
                        name = v2.decode()
                    elif n2 == 2:
                        sl = v2
                    elif n2 == 3:
                        tp = v2
                if name not in meta:
                    raise ValueError("slice of %r precedes / lacks its metadata" % name)
                shape, dtype = meta[name]
                if dtype not in DT:
                    continue
                if name not in out:
                    out[name] = np.zeros(shape, dtype=DT[dtype])
                where = _parse_slice(sl, shape)
                part = _parse_tensor_proto(tp, DT[dtype])
                target = out[name][where] if shape else out[name]
                if part.size != target.size:
                    raise ValueError("tensor %r: slice holds %d values, expected %d" % (name, part.size, target.size))
                if shape:
                    out[name][where] = part.reshape(target.shape)
                else:
                    out[name][...] = part.reshape(())
    missing = set(meta) - set(out) - {n for n, (_, d) in meta.items() if d not in DT}
    if missing:
        raise ValueError("tensors without data: %s" % sorted(missing))
    return out


def _ordered_string(s: bytes) -> bytes:
    """OrderedCode::WriteString: 0x00 -> 00 ff, 0xff -> ff 00, terminated by 00 01"""
    return bytes(b for c in s for b in ((0, 0xFF) if c == 0 else (0xFF, 0) if c == 0xFF else (c,))) + b"\x00\x01"


def _v1_key(name: str, rank: int) -> bytes:
    """EncodeTensorNameSlice for a full slice: num 0, the name, the rank, (start 0, length -1) per dimension"""
    num = lambda v: b"\x00" if v == 0 else bytes([1, v])  # OrderedCode::WriteNumIncreasing, v < 256
    return num(0) + _ordered_string(name.encode()) + num(rank) + b"\x80\x7f" * rank


def write_v1(path, arrays):
    """{name: ndarray} -> one V1 checkpoint file, every tensor as a single full slice (what tf.train.Saver wrote
    up to r0.11); floats go to float_val, integers to int_val as TensorSliceWriter::SaveData does"""
    metas, items = b"", []
    for name in sorted(arrays):
        a = np.asarray(arrays[name], order="C")  # (ascontiguousarray would turn scalars into 1-vectors)
        if a.dtype not in DT_CODE:
            raise ValueError("%s: unsupported dtype %s" % (name, a.dtype))
        code = DT_CODE[a.dtype]
        full = b"".join(_ld(1, b"") for _ in a.shape)  # Extent without start/length == the whole dimension
        metas += _ld(1, _ld(1, name.encode()) + _ld(2, _shape_msg(a.shape)) + _tag(3, 0) + _put_varint(code) + _ld(4, full))
        if code == 1:
            payload = _ld(5, a.astype("<f4").tobytes())
        elif code == 2:
            payload = _ld(6, a.astype("<f8").tobytes())
        else:
            payload = _ld(7 if code == 3 else 10, b"".join(_put_varint(int(v)) for v in a.reshape(-1)))
        tensor = _tag(1, 0) + _put_varint(code) + _ld(2, _shape_msg(a.shape)) + payload
        items.append((_v1_key(name, a.ndim), _ld(2, _ld(1, name.encode()) + _ld(2, full) + _ld(3, tensor))))
    versions = _tag(1, 0) + _put_varint(1)
    items.append((b"", _ld(1, metas + _ld(2, versions))))
    write_table(path, items)


# ------------------------------------------------------------------------------------------ front door
def find(filename):
    """'v2' / 'v1' / None for the checkpoint `tf.train.Saver.save(sess, filename)` would have left at `filename`"""
    if os.path.exists(filename + ".index"):
        return "v2"
    if os.path.isfile(filename):
        try:
            with open(filename, "rb") as f:
                f.seek(-8, os.SEEK_END)
                if struct.unpack("<Q", f.read(8))[0] == MAGIC:
                    return "v1"
        except (OSError, struct.error):
            pass
    return None


def read(filename, verify=True):
    """{variable name: ndarray} of a TensorFlow checkpoint in either format"""
    kind = find(filename)
    if kind == "v2":
        return read_v2(filename, verify)
    if kind == "v1":
        return read_v1(filename, verify)
    raise FileNotFoundError("no TensorFlow checkpoint at %s" % filename)
