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
EXACT_TRIES = 4096


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

    def _rejected(self, picks):
        rejected = np.zeros(len(picks), dtype=bool)
        seen = {}
        for j in range(len(picks)):
            if picks[j] in seen:
                rejected[j] = True
            else:
                seen[picks[j]] = j
        return rejected

    def _mass_excl(self, node, shift, held):
        """fl(tree[node] - sum of the held priorities below node), held leaves added in slot order (float32)."""
        ex = np.float32(0)
        for l in held:
            if l >= 0 and ((self.P + l) >> shift) == node:
                ex = np.float32(ex + self.tree[self.P + l])
        m = np.float32(self.tree[node] - ex)
        return m if m > 0 else np.float32(0)

    def descend_excl(self, u, held):
        """One exact draw of successive sampling without replacement: descent over the tree minus the held leaves."""
        depth = self.P.bit_length() - 1
        v = np.float32(np.float32(u) * self._mass_excl(1, depth, held))
        node = 1
        for shift in range(depth - 1, -1, -1):
            l = self._mass_excl(2 * node, shift, held)
            r = self._mass_excl(2 * node + 1, shift, held)
            left = (v < l) or (r == 0)
            if not left:
                v = np.float32(v - l)
            node = 2 * node + (0 if left else 1)
        return node - self.P

    def sample(self, B, seed, step):
        """Return (leaf indices int64[B], attempts uint32[B]) for sampling call number `step`.
        MAX_ROUNDS rounds of reject-and-redraw; slots still rejected after that are resolved exactly, in slot order, by a descent that
        leaves out the held leaves (uniforms named by attempt 0x80000000 + t) - the reference's sampler always succeeds once
        curr_size >= batch_size (StatsBase.sample(...; replace=false), src/prioritized_experience_replay.jl:83-85)."""
        slots = np.arange(B, dtype=np.uint32)
        attempts = np.zeros(B, dtype=np.uint32)
        picks = self.descend(sample_uniforms(seed, step, slots, attempts))
        for _ in range(MAX_ROUNDS):
            rejected = self._rejected(picks)
            if not rejected.any():
                return picks, attempts
            attempts[rejected] += 1
            picks[rejected] = self.descend(sample_uniforms(seed, step, slots[rejected], attempts[rejected]))
        rejected = self._rejected(picks)
        if not rejected.any():
            return picks, attempts
        held = [(-1 if rejected[j] else int(picks[j])) for j in range(B)]
        for s in range(B):
            if held[s] >= 0:
                continue
            for t in range(EXACT_TRIES):
                att = np.uint32(0x80000000 + t)
                pick = int(self.descend_excl(sample_uniforms(seed, step, np.array([s], np.uint32), np.array([att], np.uint32))[0], held))
                if pick not in held:
                    break
            else:
                raise RuntimeError("sum-tree sampling did not reach distinct indices")
            held[s] = pick
            attempts[s] = att
        return np.asarray(held, np.int64), attempts
