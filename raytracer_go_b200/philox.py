"""Vectorised Philox4x32 (Random123; Salmon et al., SC'11) in numpy.

Used only to synthesise deterministic scenes (the reference seeds its scene RNG from the clock,
main.go:246), with Random123's default 10 rounds.  The render-time streams live in the CUDA kernels
(csrc/rt_rng.h) and use 7 rounds; the scenes do not depend on that choice.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = 0x9E3779B9
_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
_S32 = np.uint64(32)


def philox4x32_10(ctr, key, rounds=10):
    """ctr: (..., 4) uint32, key: (2,) uint32-like -> (..., 4) uint32."""
    ctr = np.asarray(ctr, dtype=np.uint32)
    c0, c1, c2, c3 = (ctr[..., i].astype(np.uint64) for i in range(4))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = _M0 * c0
        p1 = _M1 * c2
        n0 = (p1 >> _S32) ^ c1 ^ np.uint64(k0)
        n1 = p1 & _MASK
        n2 = (p0 >> _S32) ^ c3 ^ np.uint64(k1)
        n3 = p0 & _MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def u32_to_f32(u):
    """rand.Float32()-like uniform on [0,1): 24 mantissa bits, exactly representable."""
    return (np.asarray(u, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)


def stream_floats(seed, a, b, n_floats, tag=0):
    """For every element of integer arrays a,b: the first n_floats floats of the stream whose
    counter is (a, b, block, tag) and key (seed_lo, seed_hi).  Returns (..., n_floats) float32."""
    a = np.asarray(a, dtype=np.uint32)
    b = np.broadcast_to(np.asarray(b, dtype=np.uint32), a.shape)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    blocks = (n_floats + 3) // 4
    out = []
    for blk in range(blocks):
        ctr = np.stack([a, b, np.full(a.shape, blk, np.uint32), np.full(a.shape, tag, np.uint32)], axis=-1)
        out.append(u32_to_f32(philox4x32_10(ctr, key)))
    return np.concatenate(out, axis=-1)[..., :n_floats]
