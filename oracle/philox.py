"""Host restatement of the DropPath generator of unite_b200/csrc/rng.cu — TEST INFRASTRUCTURE (see unite_oracle.py).

Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11; the same round function as
cuRAND / torch's CUDA generator): key = (seed_lo, seed_hi), counter = (q, 0, step_lo, step_hi) for the q-th group of four
outputs of draw number `step`.  Element e of the [depth, 2, B] factor tensor takes output e % 4 of group e // 4:
    u = (x >> 8) * 2**-24,    factor = floor(keep_l + u) / keep_l   in fp32,   keep_l = 1 - rate_l  (fp32)
which is timm 0.4.12 drop_path (src/models/modeling_finetune.py:42-50) with the uniform draw made explicit.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over numpy uint32 arrays c0..c3; scalars k0, k1 (uint32)."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = np.uint32(k0 + W0), np.uint32(k1 + W1)
    return c0, c1, c2, c3


def drop_path_factors(rates, B, seed, step):
    """-> float32 [depth, 2, B], bit-identical to ub_drop_path_draw for the same (rates, B, seed, step)."""
    rates = np.asarray(rates, dtype=np.float32)
    depth = rates.shape[0]
    n = depth * 2 * B
    q = np.arange((n + 3) // 4, dtype=np.uint32)
    z = np.zeros_like(q)
    step = int(step)
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    r = philox4x32_10(q, z, z + np.uint32(step & 0xFFFFFFFF), z + np.uint32(step >> 32), seed & 0xFFFFFFFF, seed >> 32)
    x = np.stack(r, axis=1).reshape(-1)[:n]
    u = (x >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
    keep = (np.float32(1.0) - rates)[np.arange(n) // (2 * B)]
    return (np.floor(keep + u) / keep).astype(np.float32).reshape(depth, 2, B)
