"""Host restatement of the on-device synthetic embedding generator (csrc/kernels_util.cuh
`synth_rows_kernel`): bit-identical rows on CPU and GPU, so that any row of a 38.6M-row device-
resident collection can be regenerated on the host for checking (SURVEY.md §8d "synthetic inputs").

Row r of stream (seed, stream):
    for chunk c in 0..191:  (w0,w1,w2,w3) = Philox4x32-10(counter=(r_lo, r_hi, c, 0), key=(k0, k1))
        component[4c+j] = byte-sum(w_j) - 510                  (Irwin-Hall(4) on bytes, integer)
    ss  = sum(component^2)                                      (integer, < 2^31)
    inv = fl32(norm) / sqrt_fl32(fl32(ss))                      (IEEE, correctly rounded)
    x[t] = fl32(component[t]) * inv
with k0 = lo32(seed) ^ lo32(stream), k1 = hi32(seed) ^ hi32(stream) ^ 0x5eed.
Streams used by the benchmark: passages stream 0, queries stream 1 (seed 0).

`mean_shift = M > 0` adds sign[t] * M to component[t] before the normalisation, sign[t] = +-1 from the low
bit of word t of Philox(counter=(0xffffffff, 0xffffffff, c, 0x6d65616e), key=(lo32(seed), hi32(seed) ^ 0x5eed)):
one fixed direction per seed, shared by passages and queries.  The rows then look like LayerNorm outputs
with a common mean (reference model/models.py:136-145): cos(p, p') ~ M^2 / (M^2 + 147.8^2); M = 443 -> 0.9.
"""
from __future__ import annotations

import numpy as np

DIM = 768
_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def _keys(seed: int, stream: int) -> tuple[int, int]:
    k0 = (seed & 0xFFFFFFFF) ^ (stream & 0xFFFFFFFF)
    k1 = ((seed >> 32) & 0xFFFFFFFF) ^ ((stream >> 32) & 0xFFFFFFFF) ^ 0x5EED
    return k0, k1


def philox4x32_10(c0, c1, c2, c3, k0: int, k1: int):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11).  Inputs are uint64 arrays holding 32-bit values."""
    c0 = c0.astype(np.uint64); c1 = c1.astype(np.uint64); c2 = c2.astype(np.uint64); c3 = c3.astype(np.uint64)
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
    return c0, c1, c2, c3


def _bytesum(w):
    return ((w & np.uint64(0xFF)) + ((w >> np.uint64(8)) & np.uint64(0xFF)) + ((w >> np.uint64(16)) & np.uint64(0xFF))
            + (w >> np.uint64(24))).astype(np.int64) - 510


def mean_signs(seed: int = 0) -> np.ndarray:
    """int64 [768] of +-1: the fixed direction of the common mean for this seed."""
    chunk = np.arange(DIM // 4, dtype=np.uint64)
    ones = np.full_like(chunk, 0xFFFFFFFF)
    w = philox4x32_10(ones, ones, chunk, np.full_like(chunk, 0x6D65616E), int(seed) & 0xFFFFFFFF,
                      ((int(seed) >> 32) & 0xFFFFFFFF) ^ 0x5EED)
    bits = np.stack([x & np.uint64(1) for x in w], axis=-1).reshape(DIM).astype(np.int64)
    return 2 * bits - 1


def rows(row_ids, seed: int = 0, stream: int = 0, norm: float = 1.0, mean_shift: int = 0) -> np.ndarray:
    """float32 [len(row_ids), 768]: the given rows of stream (seed, stream)."""
    row_ids = np.asarray(row_ids, dtype=np.uint64).reshape(-1)
    k0, k1 = _keys(int(seed), int(stream))
    shift = mean_signs(seed) * int(mean_shift) if mean_shift else None
    out = np.empty((row_ids.size, DIM), dtype=np.float32)
    chunk = np.arange(DIM // 4, dtype=np.uint64)[None, :]
    step = 8192
    for a in range(0, row_ids.size, step):
        r = row_ids[a:a + step][:, None]
        c0 = np.broadcast_to(r & _MASK, (r.shape[0], DIM // 4))
        c1 = np.broadcast_to(r >> np.uint64(32), (r.shape[0], DIM // 4))
        c2 = np.broadcast_to(chunk, (r.shape[0], DIM // 4))
        c3 = np.zeros_like(c2)
        w = philox4x32_10(c0, c1, c2, c3, k0, k1)
        comp = np.stack([_bytesum(x) for x in w], axis=-1).reshape(r.shape[0], DIM)  # [rows, 192, 4] -> 768
        if shift is not None:
            comp = comp + shift[None, :]
        ss = (comp * comp).sum(axis=1)
        with np.errstate(divide="ignore"):
            inv = np.float32(norm) / np.sqrt(ss.astype(np.float32))
        inv = np.where(ss > 0, inv, np.float32(0)).astype(np.float32)
        out[a:a + step] = comp.astype(np.float32) * inv[:, None]
    return out


def block(first_row: int, n: int, seed: int = 0, stream: int = 0, norm: float = 1.0, mean_shift: int = 0) -> np.ndarray:
    """float32 [n, 768]: rows first_row .. first_row+n-1."""
    return rows(np.arange(first_row, first_row + n, dtype=np.uint64), seed, stream, norm, mean_shift)
