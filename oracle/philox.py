"""Philox4x32-10 counter-based RNG (oracle side; TEST INFRASTRUCTURE).

The reference draws its sample indices from the buffer's MersenneTwister(0) through StatsBase
(src/prioritized_experience_replay.jl:42,85).  That stream cannot be reproduced offline (SURVEY
App. B.5), so the engine defines its own counter-based stream and this file restates it:

    key     = (seed_lo, seed_hi)
    counter = (slot j, attempt, step_lo, step_hi)
    u       = float32(x0 >> 8) * 2^-24            in [0, 1)

Philox4x32-10 follows Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11);
tests/test_oracle_cpu.py checks it against the Random123 known-answer vectors.
"""
import numpy as np

_M0 = np.uint64(0xD2511F53)
_M1 = np.uint64(0xCD9E8D57)
_W0 = np.uint32(0x9E3779B9)
_W1 = np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32(counter, key, rounds=10):
    """counter: (..., 4) uint32, key: (..., 2) uint32 (broadcastable) -> (..., 4) uint32."""
    c = np.asarray(counter, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    c0, c1, c2, c3 = [c[..., i].astype(np.uint64) for i in range(4)]
    k0 = np.broadcast_to(k[..., 0], c0.shape).astype(np.uint32)
    k1 = np.broadcast_to(k[..., 1], c0.shape).astype(np.uint32)
    with np.errstate(over="ignore"):
        for r in range(rounds):
            p0 = _M0 * c0
            p1 = _M1 * c2
            hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
            hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
            n0 = hi1 ^ c1 ^ k0.astype(np.uint64)
            n2 = hi0 ^ c3 ^ k1.astype(np.uint64)
            c0, c1, c2, c3 = n0, lo1, n2, lo0
            if r + 1 < rounds:
                k0 = (k0 + _W0).astype(np.uint32)
                k1 = (k1 + _W1).astype(np.uint32)
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def uniform24(x):
    """uint32 -> float32 uniform in [0,1) with 24 random bits (exactly representable)."""
    return (np.asarray(x, dtype=np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def sample_uniforms(seed, step, slots, attempts):
    """Uniforms for (slot j, attempt) pairs of sampling call number `step`."""
    slots = np.asarray(slots, dtype=np.uint32)
    attempts = np.broadcast_to(np.asarray(attempts, dtype=np.uint32), slots.shape)
    ctr = np.stack([slots, attempts,
                    np.full(slots.shape, step & 0xFFFFFFFF, dtype=np.uint32),
                    np.full(slots.shape, (step >> 32) & 0xFFFFFFFF, dtype=np.uint32)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return uniform24(philox4x32(ctr, key)[..., 0])
