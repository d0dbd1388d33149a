"""numpy restatement of tfkaldi_b200/csrc/philox.cuh (Philox4x32-10) so the oracle regenerates the
GPU's dropout masks bit-exactly.  TEST INFRASTRUCTURE ONLY.

Mask semantics: tf.nn.dropout(x, keep) = x/keep * floor(keep + u), u ~ U[0,1)
(reference: neuralNetworks/classifiers/activation.py:140-141).  One Philox call yields eight 16-bit fields f; we draw
u = f * 2^-16 and keep the element iff f >= ceil((1-keep) * 2^16), i.e. floor(keep+u) == 1 in exact integer
arithmetic.  Counter = (col >> 3, row, 0, 0), key = (seed & 0xffffffff, seed >> 32); output word i serves columns
8*(col>>3) + 2i (low half) and + 2i + 1 (high half).
"""
import math

import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised over uint32 counter arrays; returns 4 uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    k0 &= 0xFFFFFFFF
    k1 &= 0xFFFFFFFF
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return tuple(x.astype(np.uint32) for x in (c0, c1, c2, c3))


def dropout_threshold(keep: float) -> int:
    return int(math.ceil((1.0 - float(np.float32(keep))) * 65536.0))


def dropout_keep_mask(seed: int, rows: int, cols: int, keep: float) -> np.ndarray:
    """Boolean [rows, cols]: True where the unit is kept."""
    seed &= 0xFFFFFFFFFFFFFFFF
    groups = (cols + 7) // 8
    r = np.arange(rows, dtype=np.uint64)[:, None] + np.zeros((1, groups), dtype=np.uint64)
    g = np.arange(groups, dtype=np.uint64)[None, :] + np.zeros((rows, 1), dtype=np.uint64)
    zeros = np.zeros_like(r)
    w = philox4x32_10(g, r, zeros, zeros, seed & 0xFFFFFFFF, seed >> 32)
    words = np.stack(w, axis=-1)  # [rows, groups, 4]
    fields = np.stack([words & np.uint32(0xFFFF), words >> np.uint32(16)], axis=-1)  # [rows, groups, 4, 2]: low half first
    fields = fields.reshape(rows, groups * 8)[:, :cols]
    return fields >= np.uint32(dropout_threshold(keep))
