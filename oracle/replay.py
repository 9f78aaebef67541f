"""PrioritizedReplayBuffer restatement (TEST INFRASTRUCTURE).

Follows src/prioritized_experience_replay.jl line by line:
  :3-17    DQExperience            (s, a::Int32 1-based, r::Float32, sp, done::Bool)
  :39-58   constructor             defaults alpha=0.6, beta=0.4, eps=1e-3
  :65-74   add_exp!                priority (td+eps)^alpha, ring index mod1
  :76-80   update_priorities!      (|td|+eps)^alpha, must be > 0
  :82-87   sample                  StatsBase weighted sampling without replacement (A-ExpJ)
  :89-104  get_batch               gather + importance weights (n*p_i/sum p)^(-beta)

Two index sources are offered: `sample_indices_aexpj` restates the reference's O(N) sampler with
numpy's generator (Julia's MersenneTwister stream is not reproducible offline - parity of the path
is defined on identical sampled batches), and `sample_indices_sumtree` is the engine's sum-tree
(oracle/sumtree.py), bit-exact against the CUDA kernel.
"""
import heapq
import math
from dataclasses import dataclass

import numpy as np

from .sumtree import SumTree


def pow_f32(x, y):
    """Float32 ^ Float32 evaluated through Float64 and rounded once (Julia's Float32 pow)."""
    return np.power(np.asarray(x, np.float32).astype(np.float64), np.float64(np.float32(y))).astype(np.float32)


def pairwise_sum_f32(x, blk=1024):
    """Julia Base.mapreduce_impl pairwise float32 sum (blocks < 1024 summed in index order)."""
    x = np.asarray(x, np.float32)
    n = x.size
    if n == 0:
        return np.float32(0)
    if n < blk:
        return np.cumsum(x, dtype=np.float32)[-1]
    mid = (n - 1) // 2 + 1                        # imid = ifirst + ((ilast-ifirst)>>1), 1-based inclusive
    return np.float32(pairwise_sum_f32(x[:mid], blk) + pairwise_sum_f32(x[mid:], blk))


@dataclass
class DQExperience:
    s: np.ndarray
    a: int          # 1-based action index (Int32 in the reference)
    r: float
    sp: np.ndarray
    done: bool


class PrioritizedReplayBuffer:
    def __init__(self, obs_shape, max_size, batch_size, alpha=0.6, beta=0.4, eps=1e-3, seed=0, obs_dtype=np.float32):
        self.max_size, self.batch_size = int(max_size), int(batch_size)
        self.alpha, self.beta, self.eps = np.float32(alpha), np.float32(beta), np.float32(eps)
        self.rng = np.random.default_rng(seed)
        self._curr_size = 0
        self._idx = 0                                # 0-based here; the reference's _idx is this + 1
        self._priorities = np.zeros(self.max_size, np.float32)
        self.obs_shape = tuple(obs_shape)
        self._s = np.zeros((self.max_size,) + self.obs_shape, obs_dtype)
        self._sp = np.zeros((self.max_size,) + self.obs_shape, obs_dtype)
        self._a = np.zeros(self.max_size, np.int32)
        self._r = np.zeros(self.max_size, np.float32)
        self._done = np.zeros(self.max_size, np.uint8)
        self.tree = SumTree(self.max_size)

    # :65-74
    def add_exp(self, s, a, r, sp, done, td_err=None):
        td_err = abs(np.float32(r)) if td_err is None else np.float32(td_err)
        assert td_err + self.eps > 0
        prio = pow_f32(np.float32(td_err + self.eps), self.alpha)
        i = self._idx
        self._s[i], self._a[i], self._r[i], self._sp[i], self._done[i] = s, a, r, sp, done
        self._priorities[i] = prio
        self.tree.set_leaves([i], [prio])
        self._idx = (self._idx + 1) % self.max_size
        if self._curr_size < self.max_size:
            self._curr_size += 1

    def add_batch(self, s, a, r, sp, done, td_err):
        for k in range(len(a)):
            self.add_exp(s[k], a[k], r[k], sp[k], done[k], td_err[k])

    # :76-80
    def update_priorities(self, indices, td_errors):
        new = pow_f32(np.abs(np.asarray(td_errors, np.float32)) + self.eps, self.alpha)
        assert np.all(new > 0)
        self._priorities[indices] = new
        self.tree.set_leaves(indices, new)

    # :82-87, index part
    def sample_indices_aexpj(self):
        assert self._curr_size >= self.batch_size and self.max_size >= self.batch_size
        w = self._priorities[:self._curr_size].copy()      # r._priorities[1:n] allocates a copy
        _ = w.sum()                                        # Weights(...) computes the sum
        return efraimidis_aexpj_wsample_norep(self.rng, w, self.batch_size)

    def sample_indices_sumtree(self, seed, step):
        assert self._curr_size >= self.batch_size
        return self.tree.sample(self.batch_size, seed, step)[0]

    # :89-104
    def get_batch(self, idx, total="pairwise", dequant=None):
        idx = np.asarray(idx, np.int64)
        assert idx.size == self.batch_size
        s = self._s[idx]
        sp = self._sp[idx]
        if dequant is not None:
            s, sp = dequant(s), dequant(sp)
        a = self._a[idx].astype(np.int64)
        r = self._r[idx].copy()
        done = self._done[idx].astype(np.float32)
        pw = self._priorities[idx].copy()
        n = self._curr_size
        if isinstance(total, str):
            tot = pairwise_sum_f32(self._priorities[:n]) if total == "pairwise" else self.tree.total
        else:
            tot = np.float32(total)
        p = pw / np.float32(tot)
        weights = pow_f32(np.float32(n) * p, -self.beta)
        return s, a, r, sp, done, idx, weights


def efraimidis_aexpj_wsample_norep(rng, w, k):
    """StatsBase.efraimidis_aexpj_wsample_norep!; returns 0-based indices in
    descending key order.  Keys are Float64 as in StatsBase; the jump scan is done on a float64
    prefix sum (same crossing points as the reference's running subtraction up to rounding)."""
    n = w.size
    pq = []
    i = 0
    while len(pq) < k and i < n:
        if w[i] > 0:
            pq.append((float(w[i]) / rng.exponential(), i))
        i += 1
    if len(pq) < k:
        raise ValueError("not enough positive weights")
    heapq.heapify(pq)
    threshold = pq[0][0]
    c = np.cumsum(w[i:], dtype=np.float64)
    base = 0.0
    pos = 0
    m = c.size
    while pos < m:
        X = threshold * rng.exponential()
        j = int(np.searchsorted(c, base + X, side="left"))   # first j with c[j] - base >= X  <=>  X - sum <= 0
        if j >= m:
            break
        wj = float(w[i + j])
        t = math.exp(-wj / threshold)
        key = -wj / math.log(t + rng.random() * (1.0 - t))
        heapq.heapreplace(pq, (key, i + j))
        threshold = pq[0][0]
        base = c[j]
        pos = j + 1
    out = np.empty(k, np.int64)
    for q in range(k - 1, -1, -1):
        out[q] = heapq.heappop(pq)[1]
    return out
