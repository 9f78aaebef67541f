"""Engine: thin object wrapper over the C-ABI handle (one per GPU)."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import lib, ptr, vptr, DQNError


def make_config(layers, obs_shape, n_actions, *, obs_dtype="f32", dueling=True, double_q=True, prioritized_replay=True,
                batch_size=32, buffer_size=1000, alpha=0.6, beta=0.4, eps=1e-3, learning_rate=1e-4, discount=1.0,
                seed=0, device=0, math_mode=_capi.MATH_FP32, use_graph=True, rank=0, world=1, nccl_id=None, max_act_rows=0,
                trace_length=0, max_episode_length=0):
    """obs_shape is the Flux size tuple of one observation: (d,) or (W, H, C)."""
    cfg = _capi.default_config()
    if len(obs_shape) == 1:
        cfg.obs_c, cfg.obs_h, cfg.obs_w = int(obs_shape[0]), 1, 1
    elif len(obs_shape) == 3:
        cfg.obs_w, cfg.obs_h, cfg.obs_c = (int(x) for x in obs_shape)
    elif len(obs_shape) == 2:
        cfg.obs_w, cfg.obs_h, cfg.obs_c = int(obs_shape[0]), int(obs_shape[1]), 1
    else:
        raise ValueError("observation must have 1 to 3 dimensions")
    cfg.obs_dtype = _capi.OBS_U8 if obs_dtype in ("u8", np.uint8, _capi.OBS_U8) else _capi.OBS_F32
    cfg.n_actions = int(n_actions)
    descs = [l.desc() if hasattr(l, "desc") else l for l in layers]
    if len(descs) > _capi.DQN_MAX_LAYERS:
        raise ValueError("too many layers")
    cfg.n_layers = len(descs)
    for i, d in enumerate(descs):
        for k, v in d.items():
            setattr(cfg.layers[i], k, int(v))
    cfg.dueling, cfg.double_q, cfg.prioritized_replay = int(dueling), int(double_q), int(prioritized_replay)
    cfg.batch_size, cfg.buffer_size = int(batch_size), int(buffer_size)
    cfg.alpha, cfg.beta, cfg.eps = float(alpha), float(beta), float(eps)
    cfg.learning_rate, cfg.discount = float(learning_rate), float(discount)
    cfg.seed, cfg.device, cfg.math_mode, cfg.use_graph = int(seed), int(device), int(math_mode), int(use_graph)
    cfg.rank, cfg.world, cfg.max_act_rows = int(rank), int(world), int(max_act_rows)
    cfg.trace_length, cfg.max_episode_length = int(trace_length), int(max_episode_length)
    if nccl_id is not None:
        C.memmove(cfg.nccl_id, bytes(nccl_id), _capi.DQN_NCCL_ID_BYTES)
    return cfg


def nccl_unique_id():
    buf = (C.c_uint8 * _capi.DQN_NCCL_ID_BYTES)()
    rc = lib.dqn_nccl_unique_id(buf)
    if rc != 0:
        raise DQNError(rc, (lib.dqn_last_error(None) or b"").decode())
    return bytes(buf)


class Engine:
    def __init__(self, cfg, handle=None):
        """handle: wrap an engine owned by someone else (a rank of a Group) instead of creating one"""
        self.cfg = cfg
        self._owned = handle is None
        if handle is None:
            h = C.c_void_p()
            rc = lib.dqn_engine_create(C.byref(cfg), C.byref(h))
            if rc != 0:
                raise DQNError(rc, (lib.dqn_last_error(None) or b"").decode())
        else:
            h = C.c_void_p(handle)
        self.h = h
        self.recurrent = any(cfg.layers[i].kind == _capi.LAYER_LSTM for i in range(cfg.n_layers))
        self.T = (cfg.trace_length or 40) if self.recurrent else 1
        self.Bep = cfg.batch_size
        self.B = cfg.batch_size * self.T          # rows of the step's diagnostics (recurrent: trace_length * batch_size, time-major)
        self.nA = cfg.n_actions
        self.obs_elems = cfg.obs_c * cfg.obs_h * cfg.obs_w
        self.obs_np = np.uint8 if cfg.obs_dtype == _capi.OBS_U8 else np.float32
        self.obs_shape = (cfg.obs_c, cfg.obs_h, cfg.obs_w) if (cfg.obs_h, cfg.obs_w) != (1, 1) else (cfg.obs_c,)
        self.num_params = int(lib.dqn_num_params(h))

    def close(self):
        if getattr(self, "h", None):
            if self._owned:
                lib.dqn_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise DQNError(rc, (lib.dqn_last_error(self.h) or b"").decode())

    # ---- parameters --------------------------------------------------------------------------------
    def set_params(self, flat, which=_capi.NET_ONLINE):
        flat = np.ascontiguousarray(flat, np.float32)
        self._ck(lib.dqn_set_params(self.h, which, ptr(flat, C.c_float), flat.size))

    def get_params(self, which=_capi.NET_ONLINE):
        out = np.empty(self.num_params, np.float32)
        self._ck(lib.dqn_get_params(self.h, which, ptr(out, C.c_float), out.size))
        return out

    def sync_target(self):
        self._ck(lib.dqn_sync_target(self.h))

    def get_adam_state(self):
        m = np.empty(self.num_params, np.float32)
        v = np.empty(self.num_params, np.float32)
        bp = (C.c_double * 2)()
        self._ck(lib.dqn_get_adam_state(self.h, ptr(m, C.c_float), ptr(v, C.c_float), bp, m.size))
        return m, v, (bp[0], bp[1])

    # ---- replay ------------------------------------------------------------------------------------
    def _obs(self, x, n):
        x = np.ascontiguousarray(x, self.obs_np)
        if x.size != n * self.obs_elems:
            raise ValueError(f"observations: expected {n}x{self.obs_elems} elements, got {x.size}")
        return x

    def replay_add(self, s, a, r, sp, done, td0):
        a = np.ascontiguousarray(a, np.int32)
        n = a.size
        s, sp = self._obs(s, n), self._obs(sp, n)
        r = np.ascontiguousarray(r, np.float32)
        done = np.ascontiguousarray(done, np.uint8)
        td0 = np.ascontiguousarray(td0, np.float32)
        assert r.size == n and done.size == n and td0.size == n
        self._ck(lib.dqn_replay_add(self.h, vptr(s), ptr(a, C.c_int32), ptr(r, C.c_float), vptr(sp), ptr(done, C.c_uint8), ptr(td0, C.c_float), n))

    def replay_size(self):
        n, c = C.c_int64(), C.c_int64()
        self._ck(lib.dqn_replay_size(self.h, C.byref(n), C.byref(c)))
        return n.value, c.value

    def replay_fill_synthetic(self, n, seed):
        self._ck(lib.dqn_replay_fill_synthetic(self.h, int(n), int(seed)))

    def replay_read(self, idx):
        idx = np.ascontiguousarray(idx, np.int64)
        n = idx.size
        s = np.empty((n,) + self.obs_shape, self.obs_np)
        sp = np.empty_like(s)
        a = np.empty(n, np.int32)
        r = np.empty(n, np.float32)
        d = np.empty(n, np.uint8)
        self._ck(lib.dqn_replay_read(self.h, ptr(idx, C.c_int64), n, vptr(s), ptr(a, C.c_int32), ptr(r, C.c_float), vptr(sp), ptr(d, C.c_uint8)))
        return s, a, r, sp, d

    def update_priorities(self, idx, td):
        idx = np.ascontiguousarray(idx, np.int64)
        td = np.ascontiguousarray(td, np.float32)
        self._ck(lib.dqn_update_priorities(self.h, ptr(idx, C.c_int64), ptr(td, C.c_float), idx.size))

    def set_priorities(self, idx, prio):
        idx = np.ascontiguousarray(idx, np.int64)
        prio = np.ascontiguousarray(prio, np.float32)
        self._ck(lib.dqn_set_priorities(self.h, ptr(idx, C.c_int64), ptr(prio, C.c_float), idx.size))

    def get_priorities(self, n=None):
        n = self.replay_size()[0] if n is None else n
        out = np.empty(n, np.float32)
        self._ck(lib.dqn_get_priorities(self.h, ptr(out, C.c_float), n))
        return out

    def get_tree(self):
        out = np.empty(int(lib.dqn_tree_nodes(self.h)), np.float32)
        self._ck(lib.dqn_get_tree(self.h, ptr(out, C.c_float), out.size))
        return out

    # ---- EpisodeReplayBuffer (recurrent engines) ---------------------------------------------------
    def episode_add(self, s, a, r, sp, done):
        a = np.ascontiguousarray(a, np.int32)
        n = a.size
        s = np.ascontiguousarray(s, np.float32); sp = np.ascontiguousarray(sp, np.float32)
        r = np.ascontiguousarray(r, np.float32); done = np.ascontiguousarray(done, np.uint8)
        assert s.size == n * self.obs_elems and sp.size == s.size and r.size == n and done.size == n
        self._ck(lib.dqn_episode_add(self.h, ptr(s, C.c_float), ptr(a, C.c_int32), ptr(r, C.c_float), ptr(sp, C.c_float), ptr(done, C.c_uint8), n))

    def episode_count(self):
        n, c = C.c_int64(), C.c_int64()
        self._ck(lib.dqn_episode_count(self.h, C.byref(n), C.byref(c)))
        return n.value, c.value

    def episode_sample(self, call):
        idx = np.empty(self.Bep, np.int64); start = np.empty(self.Bep, np.int32)
        self._ck(lib.dqn_episode_sample(self.h, int(call), ptr(idx, C.c_int64), ptr(start, C.c_int32)))
        return idx, start

    def policy_reset(self):
        self._ck(lib.dqn_policy_reset(self.h))

    def sample_indices(self, call):
        out = np.empty(self.B, np.int64)
        self._ck(lib.dqn_sample_indices(self.h, int(call), ptr(out, C.c_int64)))
        return out

    def get_batch(self, idx):
        idx = np.ascontiguousarray(idx, np.int64)
        assert idx.size == self.B
        s = np.empty((self.B,) + self.obs_shape, np.float32)
        sp = np.empty_like(s)
        a = np.empty(self.B, np.int32)
        r = np.empty(self.B, np.float32)
        d = np.empty(self.B, np.float32)
        w = np.empty(self.B, np.float32)
        self._ck(lib.dqn_get_batch(self.h, ptr(idx, C.c_int64), ptr(s, C.c_float), ptr(a, C.c_int32), ptr(r, C.c_float),
                                   ptr(sp, C.c_float), ptr(d, C.c_float), ptr(w, C.c_float)))
        return s, a, r, sp, d, idx, w

    # ---- the step ----------------------------------------------------------------------------------
    def train_step(self):
        loss, gn = C.c_float(), C.c_float()
        self._ck(lib.dqn_train_step(self.h, C.byref(loss), C.byref(gn)))
        return loss.value, gn.value

    def train_step_with_indices(self, idx):
        idx = np.ascontiguousarray(idx, np.int64)
        assert idx.size == (self.Bep if self.recurrent else self.B)
        loss, gn = C.c_float(), C.c_float()
        self._ck(lib.dqn_train_step_with_indices(self.h, ptr(idx, C.c_int64), C.byref(loss), C.byref(gn)))
        return loss.value, gn.value

    def train_step_async(self):
        self._ck(lib.dqn_train_step_async(self.h))

    def sync(self):
        loss, gn = C.c_float(), C.c_float()
        self._ck(lib.dqn_sync(self.h, C.byref(loss), C.byref(gn)))
        return loss.value, gn.value

    def step_result(self, back=0):
        """(loss, grad_norm) of the step launched `back` (0 or 1) steps before the latest; back=1 does not wait for the latest."""
        loss, gn = C.c_float(), C.c_float()
        self._ck(lib.dqn_step_result(self.h, back, C.byref(loss), C.byref(gn)))
        return loss.value, gn.value

    def q_values(self, obs, which=_capi.NET_ONLINE):
        obs = np.ascontiguousarray(obs, self.obs_np)
        n = obs.size // self.obs_elems
        assert n * self.obs_elems == obs.size
        out = np.empty((n, self.nA), np.float32)
        self._ck(lib.dqn_q_values(self.h, which, vptr(obs), n, ptr(out, C.c_float)))
        return out

    def act(self, obs, eps=0.0, call=0, want_q=False):
        """epsilon-greedy actions (1-based) for a batch of observations: one forward + argmax + exploration draw on the device"""
        obs = np.ascontiguousarray(obs, self.obs_np)
        n = obs.size // self.obs_elems
        a = np.empty(n, np.int32)
        q = np.empty((n, self.nA), np.float32) if want_q else None
        self._ck(lib.dqn_act(self.h, vptr(obs), n, float(eps), int(call), ptr(a, C.c_int32), ptr(q, C.c_float) if want_q else None))
        return (a, q) if want_q else a

    # ---- diagnostics -------------------------------------------------------------------------------
    def last_indices(self):
        out = np.empty(self.B if not self.recurrent else self.B, np.int64)
        self._ck(lib.dqn_get_last_indices(self.h, ptr(out, C.c_int64)))
        return out

    def td(self):
        out = np.empty(self.B, np.float32)
        self._ck(lib.dqn_get_td(self.h, ptr(out, C.c_float)))
        return out

    def is_weights(self):
        out = np.empty(self.B, np.float32)
        self._ck(lib.dqn_get_is_weights(self.h, ptr(out, C.c_float)))
        return out

    def q(self, which):
        out = np.empty((self.B, self.nA), np.float32)
        self._ck(lib.dqn_get_q(self.h, which, ptr(out, C.c_float)))
        return out

    def targets(self):
        y = np.empty(self.B, np.float32)
        b = np.empty(self.B, np.int32)
        self._ck(lib.dqn_get_targets(self.h, ptr(y, C.c_float), ptr(b, C.c_int32)))
        return y, b

    def grads(self):
        out = np.empty(self.num_params, np.float32)
        self._ck(lib.dqn_get_grads(self.h, ptr(out, C.c_float), out.size))
        return out

    def activation(self, stage, tower, shape):
        """online-network output of layer `stage` (convs first, then the tower's Dense layers) on the s rows of the last step"""
        out = np.empty((self.B,) + tuple(shape), np.float32)
        self._ck(lib.dqn_get_activation(self.h, int(stage), int(tower), ptr(out, C.c_float), out.size))
        return out

    # ---- measurement -------------------------------------------------------------------------------
    def timer_start(self):
        self._ck(lib.dqn_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        self._ck(lib.dqn_timer_stop(self.h, C.byref(ms)))
        return ms.value

    def launches_per_step(self):
        return int(lib.dqn_launches_per_step(self.h))

    def collective_kind(self):
        """0: none (world 1), 1: NCCL, 2: the engine's own all-reduce over NVLink peer memory"""
        return int(lib.dqn_collective_kind(self.h))

    def set_profiling(self, on):
        self._ck(lib.dqn_set_profiling(self.h, int(on)))

    def get_profile(self):
        buf = C.create_string_buffer(1 << 16)
        self._ck(lib.dqn_get_profile(self.h, buf, len(buf)))
        out = []
        for line in buf.value.decode().splitlines():
            name, ms, cnt, by, fl = line.split()
            out.append(dict(name=name, ms=float(ms), count=int(cnt), bytes=float(by), flops=float(fl)))
        return out

    def flush_l2(self):
        self._ck(lib.dqn_flush_l2(self.h))


class Group:
    """ndev engines in ONE process behind dqn_group_create (what a single-process host such as the reference's Julia solver binds):
    rank r on device r, one NCCL communicator inside.  `engines[r]` is an Engine view of rank r for the per-shard calls."""

    def __init__(self, cfg, ndev, devices=None):
        g = C.c_void_p()
        dev = None if devices is None else (C.c_int * ndev)(*devices)
        rc = lib.dqn_group_create(C.byref(cfg), int(ndev), dev, C.byref(g))
        if rc != 0:
            raise DQNError(rc, (lib.dqn_group_last_error(None) or b"").decode())
        self.g = g
        self.engines = [Engine(cfg, handle=lib.dqn_group_engine(g, r)) for r in range(ndev)]

    def _ck(self, rc):
        if rc != 0:
            raise DQNError(rc, (lib.dqn_group_last_error(self.g) or b"").decode())

    def set_params(self, flat, which=_capi.NET_ONLINE):
        flat = np.ascontiguousarray(flat, np.float32)
        self._ck(lib.dqn_group_set_params(self.g, which, ptr(flat, C.c_float), flat.size))

    def sync_target(self):
        self._ck(lib.dqn_group_sync_target(self.g))

    def train_step(self):
        loss, gn = C.c_float(), C.c_float()
        self._ck(lib.dqn_group_train_step(self.g, C.byref(loss), C.byref(gn)))
        return loss.value, gn.value

    def close(self):
        if getattr(self, "g", None):
            for e in self.engines:
                e.h = None
            lib.dqn_group_destroy(self.g)
            self.g = None

    __del__ = close
