"""CPU regeneration of the engine's synthetic replay fill (TEST INFRASTRUCTURE).

`dqn_replay_fill_synthetic` makes transition i a pure function of (seed, i) (uint8 observations
i.i.d. uniform, a ~ U{1..|A|}, r ~ U(-1,1), done ~ Bernoulli(0.01), priority (|r|+eps)^alpha as add_exp!
gives new transitions, src/solver.jl:92).  This file restates that generator so the oracle can rebuild any
subset of a 1M-transition buffer without holding it."""
import numpy as np

from .philox import philox4x32
from .replay import pow_f32


def _blocks(seed, i, nwords, stream):
    i = np.asarray(i, np.uint64).reshape(-1, 1)
    w = np.arange(nwords, dtype=np.uint64).reshape(1, -1)
    ctr = np.stack([np.broadcast_to(w, (i.shape[0], nwords)).astype(np.uint32),
                    np.broadcast_to(i & np.uint64(0xFFFFFFFF), (i.shape[0], nwords)).astype(np.uint32),
                    np.broadcast_to(i >> np.uint64(32), (i.shape[0], nwords)).astype(np.uint32),
                    np.full((i.shape[0], nwords), stream, np.uint32)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    return philox4x32(ctr, key)          # (n, nwords, 4) uint32


def synthetic_meta(seed, idx, n_actions, alpha=0.6, eps=1e-3):
    """(a, r, done, prio) of transitions `idx` without their observations (cheap for all 1M transitions of a shard)."""
    idx = np.asarray(idx, np.int64)
    i64 = idx.astype(np.uint64)
    ctr = np.stack([np.full(idx.size, 0xFFFFFFFF, np.uint32), (i64 & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                    (i64 >> np.uint64(32)).astype(np.uint32), np.full(idx.size, 2, np.uint32)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    c = philox4x32(ctr, key)
    a = (1 + (c[:, 0] % np.uint32(n_actions))).astype(np.int32)
    r = (((c[:, 1] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)) * np.float32(2) - np.float32(1)).astype(np.float32)
    done = ((c[:, 2] >> np.uint32(8)) < np.uint32(167772)).astype(np.uint8)
    prio = pow_f32(np.abs(r) + np.float32(eps), alpha)
    return a, r, done, prio


def synthetic_transitions(seed, idx, obs_shape, obs_u8, n_actions, alpha=0.6, eps=1e-3, hwc=None):
    """Returns (s, a, r, sp, done, prio) for transition indices `idx`, observations in Flux layout
    (numpy (n, C, H, W) or (n, d)).  `hwc`: the engine stores H,W,C when the network has a conv trunk."""
    idx = np.asarray(idx, np.int64)
    elems = int(np.prod(obs_shape))
    hwc = (len(obs_shape) == 3) if hwc is None else hwc
    out = []
    for stream in (0, 1):
        if obs_u8:
            blk = _blocks(seed, idx, (elems + 15) // 16, stream)
            flat = blk.reshape(idx.size, -1).view(np.uint8)[:, :elems]       # little-endian bytes of c0..c3
        else:
            blk = _blocks(seed, idx, (elems + 3) // 4, stream)
            u = (blk.reshape(idx.size, -1) >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)
            flat = (u * np.float32(2) - np.float32(1))[:, :elems]
        if hwc and len(obs_shape) == 3:
            c, h, w = obs_shape
            flat = flat.reshape(idx.size, h, w, c).transpose(0, 3, 1, 2)
        out.append(np.ascontiguousarray(flat.reshape((idx.size,) + tuple(obs_shape))))
    i64 = idx.astype(np.uint64)
    ctr = np.stack([np.full(idx.size, 0xFFFFFFFF, np.uint32), (i64 & np.uint64(0xFFFFFFFF)).astype(np.uint32),
                    (i64 >> np.uint64(32)).astype(np.uint32), np.full(idx.size, 2, np.uint32)], axis=-1)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], np.uint32)
    c = philox4x32(ctr, key)
    a = (1 + (c[:, 0] % np.uint32(n_actions))).astype(np.int32)
    r = ((c[:, 1] >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)) * np.float32(2) - np.float32(1)
    done = ((c[:, 2] >> np.uint32(8)) < np.uint32(167772)).astype(np.uint8)
    prio = pow_f32(np.abs(r) + np.float32(eps), alpha)
    return out[0], a, r.astype(np.float32), out[1], done, prio
