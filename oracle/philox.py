"""Philox4x32-10 in NumPy (oracle twin of csrc/philox.cuh; test infrastructure only).

Counter = (id_lo, id_hi, stage, (sweep<<8)|slot), key = (seed_lo, seed_hi).  Normals are Box-Muller
in FP64 from one 4x32 block: u1 = 1 - u53(x,y), u2 = u53(z,w), z0 = r cos(2 pi u2), z1 = r sin(2 pi u2).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
SLOT_UNIFORM = 255


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32).copy() for c in (c0, c1, c2, c3))
    k0, k1 = np.uint32(k0), np.uint32(k1)
    mask = np.uint64(0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & mask).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & mask).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def u53(hi, lo):
    v = ((hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)) >> np.uint64(11)
    return v.astype(np.float64) * (1.0 / 9007199254740992.0)


def _draw(seed, ids, stage, sweep, slot):
    ids = np.asarray(ids, dtype=np.uint64)
    n = ids.shape[0]
    c0 = (ids & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    c1 = (ids >> np.uint64(32)).astype(np.uint32)
    c2 = np.full(n, stage & 0xFFFFFFFF, dtype=np.uint32)
    c3 = np.full(n, ((sweep << 8) | slot) & 0xFFFFFFFF, dtype=np.uint32)
    return philox4x32_10(c0, c1, c2, c3, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)


def uniforms(seed, ids, stage, sweep, slot=SLOT_UNIFORM):
    x, y, _, _ = _draw(seed, ids, stage, sweep, slot)
    return u53(x, y)


def normals(seed, ids, stage, sweep, d):
    ids = np.asarray(ids, dtype=np.uint64)
    Z = np.empty((ids.shape[0], d), dtype=np.float64)
    for j in range(0, d, 2):
        x, y, z, w = _draw(seed, ids, stage, sweep, j >> 1)
        u1 = 1.0 - u53(x, y)
        u2 = u53(z, w)
        r = np.sqrt(-2.0 * np.log(u1))
        Z[:, j] = r * np.cos(2.0 * np.pi * u2)
        if j + 1 < d:
            Z[:, j + 1] = r * np.sin(2.0 * np.pi * u2)
    return Z


def uniform_box(seed, ids, low, high):
    low, high = np.asarray(low, dtype=np.float64), np.asarray(high, dtype=np.float64)
    out = np.empty((len(ids), low.shape[0]), dtype=np.float64)
    for k in range(low.shape[0]):
        x, y, _, _ = _draw(seed, ids, 0xFFFFFFFF, 0, 0) if False else _draw_box(seed, ids, k)
        out[:, k] = low[k] + (high[k] - low[k]) * u53(x, y)
    return out


def _draw_box(seed, ids, k):
    # sample_box_kernel: philox_uniform(seed, id, stage=0xFFFFFFFF, sweep=k, slot=0)
    return _draw(seed, ids, 0xFFFFFFFF, k, 0)
