"""Philox4x32-10 counter-based RNG, NumPy restatement.

TEST INFRASTRUCTURE (see oracle/README.md).  The CUDA engine draws every per-shot random
number (measurement outcomes, sampled Pauli errors in ket mode) from Philox4x32-10 keyed by
``(seed)`` with counter ``(global_shot_lo, global_shot_hi, draw_index, 0)`` so that results
do not depend on how shots are sharded over GPUs.  This file restates
the published algorithm (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3",
SC'11) so the oracle can reproduce the same outcomes bit for bit.  The reference itself has
no seeded sampling test ("Nothing tests ... seeds, histograms"), so the
stream definition is ours; known-answer vectors from the Random123 distribution pin the
implementation (tests/test_philox.py).
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: (..., 4) uint32, key: (2,) or (..., 2) uint32 -> (..., 4) uint32."""
    c = np.array(counter, dtype=np.uint32, copy=True)
    k = np.array(np.broadcast_to(np.asarray(key, dtype=np.uint32), c.shape[:-1] + (2,)), copy=True)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c[..., 0].astype(np.uint64)
            p1 = _M1 * c[..., 2].astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32)
            lo0 = (p0 & _MASK).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32)
            lo1 = (p1 & _MASK).astype(np.uint32)
            n0 = hi1 ^ c[..., 1] ^ k[..., 0]
            n1 = lo1
            n2 = hi0 ^ c[..., 3] ^ k[..., 1]
            n3 = lo0
            c = np.stack([n0, n1, n2, n3], axis=-1)
            k[..., 0] = k[..., 0] + _W0
            k[..., 1] = k[..., 1] + _W1
    return c


def uniform53(words_hi, words_lo):
    """Two uint32 words -> double in [0, 1) with 53 random bits (same formula as the kernel)."""
    a = (words_hi >> np.uint32(5)).astype(np.float64)
    b = (words_lo >> np.uint32(6)).astype(np.float64)
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0)


def shot_uniforms(seed, shot_ids, draw_index):
    """Two uniforms per shot for one stochastic op.  shot_ids: int64 array of GLOBAL shot ids."""
    shot_ids = np.asarray(shot_ids, dtype=np.uint64)
    ctr = np.zeros(shot_ids.shape + (4,), dtype=np.uint32)
    ctr[..., 0] = (shot_ids & _MASK).astype(np.uint32)
    ctr[..., 1] = (shot_ids >> np.uint64(32)).astype(np.uint32)
    ctr[..., 2] = np.uint32(draw_index & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32((draw_index >> 32) & 0xFFFFFFFF)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    key = np.array([seed & 0xFFFFFFFF, seed >> 32], dtype=np.uint32)
    r = philox4x32_10(ctr, key)
    return uniform53(r[..., 0], r[..., 1]), uniform53(r[..., 2], r[..., 3])
