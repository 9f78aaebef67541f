"""Sum-tree prioritized sampler (oracle side; TEST INFRASTRUCTURE).

The reference has no sum-tree: priorities are a flat Vector{Float32} and `sample` is an O(N)
weighted draw without replacement (src/prioritized_experience_replay.jl:28,82-87).  The engine
replaces it by a binary sum-tree over the same priorities (north_star).  This file restates the
engine's tree so that leaf indices can be compared bit-exactly given the same uniforms:

  * leaves live at tree[P + i], P = capacity rounded up to a power of two, unused leaves are 0;
  * every internal node is the float32 sum  tree[k] = fl(tree[2k] + tree[2k+1])  (recomputed
    bottom-up, never updated by a delta, so the tree is a pure function of the leaves);
  * a draw descends from the root with v = fl(u * tree[1]):  go left if v < left or the right
    subtree is empty, else v = fl(v - left) and go right;
  * a batch is B draws without replacement, the semantics of StatsBase.sample(...; replace=false)
    at src/prioritized_experience_replay.jl:85 (successive draws proportional to priority):
    slot j redraws (attempt += 1) while an equal leaf is held by a slot i < j; all slots are
    checked once per round until no slot is rejected.
"""
import numpy as np

from .philox import sample_uniforms

MAX_ROUNDS = 64


def next_pow2(n):
    p = 1
    while p < n:
        p *= 2
    return p


class SumTree:
    def __init__(self, capacity):
        self.capacity = int(capacity)
        self.P = next_pow2(max(self.capacity, 2))
        self.tree = np.zeros(2 * self.P, dtype=np.float32)

    # -- construction ------------------------------------------------------------------------
    def set_leaves(self, idx, prio):
        idx = np.asarray(idx, dtype=np.int64)
        self.tree[self.P + idx] = np.asarray(prio, dtype=np.float32)
        nodes = np.unique((self.P + idx) >> 1)
        while nodes.size and nodes[0] >= 1:
            self.tree[nodes] = self.tree[2 * nodes] + self.tree[2 * nodes + 1]  # float32 add
            if nodes[0] == 1:
                break
            nodes = np.unique(nodes >> 1)

    def rebuild(self):
        lvl = self.P
        while lvl > 1:
            half = lvl // 2
            self.tree[half:lvl] = self.tree[lvl:2 * lvl:2] + self.tree[lvl + 1:2 * lvl:2]
            lvl = half

    @property
    def total(self):
        return self.tree[1]

    def leaves(self, n=None):
        n = self.capacity if n is None else n
        return self.tree[self.P:self.P + n]

    # -- sampling ----------------------------------------------------------------------------
    def descend(self, u):
        """u: float32 array in [0,1) -> leaf indices (int64)."""
        t = self.tree
        v = (np.asarray(u, dtype=np.float32) * t[1]).astype(np.float32)
        node = np.ones(v.shape, dtype=np.int64)
        for _ in range(self.P.bit_length() - 1):
            left = t[2 * node]
            right = t[2 * node + 1]
            go_left = (v < left) | (right == 0)
            v = np.where(go_left, v, (v - left).astype(np.float32))
            node = np.where(go_left, 2 * node, 2 * node + 1)
        return node - self.P

    def sample(self, B, seed, step):
        """Return (leaf indices int64[B], attempts uint32[B]) for sampling call number `step`."""
        slots = np.arange(B, dtype=np.uint32)
        attempts = np.zeros(B, dtype=np.uint32)
        picks = self.descend(sample_uniforms(seed, step, slots, attempts))
        for _ in range(MAX_ROUNDS):
            rejected = np.zeros(B, dtype=bool)
            seen = {}
            for j in range(B):
                if picks[j] in seen:
                    rejected[j] = True
                else:
                    seen[picks[j]] = j
            if not rejected.any():
                return picks, attempts
            attempts[rejected] += 1
            picks[rejected] = self.descend(sample_uniforms(seed, step, slots[rejected], attempts[rejected]))
        raise RuntimeError("sum-tree sampling did not converge to distinct indices")
