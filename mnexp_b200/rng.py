"""numpy replica of the device dropout RNG (csrc/common.cuh: lowbias32 / rng_u32).

Lets tests replay the exact GPU dropout masks into the oracle (the reference's
TF RNG cannot be matched — SURVEY.md §7 — so dropout parity is defined on our
own counter-based stream)."""
import numpy as np


def lowbias32(x):
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7feb352d)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846ca68b)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def rng_u32(seed, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    lo = (idx & np.uint64(0xffffffff)).astype(np.uint32)
    hi = (idx >> np.uint64(32)).astype(np.uint32)
    k = (np.uint64(0x9e3779b9) * np.uint64((int(seed) + 1) & 0xffffffff)) & np.uint64(0xffffffff)
    with np.errstate(over='ignore'):
        inner = lowbias32((hi + np.uint32(k)).astype(np.uint32))
    return lowbias32(lo ^ inner)


def dropout_multiplier(seed, n, p, dtype=np.float64):
    """Inverted-dropout multipliers for flat element indices 0..n-1."""
    if p <= 0:
        return np.ones(n, dtype=dtype)
    thr = np.uint32(int(np.float32(p) * np.float32(16777216.0)))
    u = rng_u32(seed, np.arange(n, dtype=np.uint64))
    keep = (u >> np.uint32(8)) >= thr
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    return np.where(keep, dtype(inv), dtype(0))


def _mulfold(a, m):
    p = a.astype(np.uint64) * np.uint64(m)
    return ((p & np.uint64(0xffffffff)) ^ (p >> np.uint64(32))).astype(np.uint32)


def quad_keep(seed, n, p):
    """Keep flags (bool, n) of the tensor-core path's dropout stream (csrc/common.cuh: quad_hash): one hash per four
    consecutive elements, 15-bit fields compared with floor(p * 32768).  n must be a multiple of 4."""
    assert n % 4 == 0
    q = np.arange(n // 4, dtype=np.uint64)
    lo = (q & np.uint64(0xffffffff)).astype(np.uint32)
    hi = (q >> np.uint64(32)).astype(np.uint32)
    k = (np.uint64(0x9e3779b9) * np.uint64((int(seed) + 1) & 0xffffffff)) & np.uint64(0xffffffff)
    with np.errstate(over='ignore'):
        key = lowbias32((hi + np.uint32(k)).astype(np.uint32))
    b = _mulfold(lo ^ key, 0x9E3779B1)
    u0 = _mulfold(b ^ np.uint32(0x85EBCA6B), 0xC2B2AE35)
    u1 = _mulfold(b ^ np.uint32(0x27D4EB2F), 0x165667B1)
    thr = np.uint32(int(np.float32(p) * np.float32(32768.0)))
    f = np.stack([u0 & np.uint32(0x7fff), (u0 >> np.uint32(16)) & np.uint32(0x7fff),
                  u1 & np.uint32(0x7fff), (u1 >> np.uint32(16)) & np.uint32(0x7fff)], -1)
    return (f >= thr).reshape(-1)


def tc_dropout_multipliers(seed, tokens, E, Ep, F, p, dtype=np.float64):
    """Inverted-dropout multipliers of the ENGINE's tensor-core path for a forward with step seed `seed`:
    X stream (seed*2) over rows padded to Ep columns, C stream (seed*2+1) over F columns, both indexed by the COMPACTED
    title index — the engine runs the title encoder over the live titles only (csrc/gather.cu: lstur_compact_titles), so
    title n with a non-zero token is row rank(n) of the streams.  tokens (N, L) -> (mx (N, L, E), mc (N, L, F)); all-pad
    titles get ones (they are never read)."""
    tokens = np.asarray(tokens)
    N, L = tokens.shape
    live = (tokens != 0).any(-1)
    n_live = int(live.sum())
    inv = np.float32(1.0) / (np.float32(1.0) - np.float32(p))
    mx, mc = np.ones((N, L, E), dtype=dtype), np.ones((N, L, F), dtype=dtype)
    if p <= 0 or n_live == 0:
        return mx, mc
    kx = quad_keep(seed * 2, n_live * L * Ep, p).reshape(n_live, L, Ep)[:, :, :E]
    kc = quad_keep(seed * 2 + 1, n_live * L * F, p).reshape(n_live, L, F)
    mx[live] = np.where(kx, dtype(inv), dtype(0))
    mc[live] = np.where(kc, dtype(inv), dtype(0))
    return mx, mc
