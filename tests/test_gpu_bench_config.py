"""GPU parity at the BENCHMARKED configuration (BASELINE.json configs[2] exactly as bench.py runs it): a 1 000 000-transition
shard filled by dqn_replay_fill_synthetic, the 2^20-leaf sum-tree, batch 256, the Nature-DQN conv network with dueling heads.

The oracle never holds the 56 GB store: oracle/synthetic.py regenerates any transition from (seed, index), and the sum-tree is
rebuilt on the CPU from the 1 M regenerated priorities.  Checked here, in both math modes:
  * every leaf priority and every one of the 2^21 tree nodes: bit-exact
  * sampled leaf indices of several sampling calls (20-level descent, shared-memory tree top): bit-exact
  * get_batch rows gathered out of the 56 GB store vs the regenerated transitions: bit-exact (k/255f0); IS weights <= 1e-6
  * full batch_train! steps (Q, targets, td, loss, gradients through the engine's ReLU masks, Adam, priority write-back and the
    refreshed tree) against oracle.forward_backward: the tolerances of tests/test_gpu_parity.py
"""
import numpy as np
import pytest

import oracle as O
import util
from oracle.synthetic import synthetic_meta, synthetic_transitions
from test_gpu_parity import check_step, SEED

pytestmark = pytest.mark.gpu

N = 1_000_000
REPLAY_SEED = 1000          # bench.py: shard_seeds(0, rank=0)["replay"]


class SyntheticShard:
    """The oracle's view of a synthetic shard: priorities + sum-tree in memory, transitions regenerated on demand."""

    def __init__(self, spec, n, seed, alpha=0.6, beta=0.4, eps=1e-3):
        self.spec, self.n, self.seed = spec, n, seed
        self.alpha, self.beta, self.eps = np.float32(alpha), np.float32(beta), np.float32(eps)
        self._a, self._r, self._done, prio = synthetic_meta(seed, np.arange(n), spec["nA"], alpha, eps)
        self._priorities = prio.copy()
        self._curr_size = n
        self.tree = O.SumTree(n)
        self.tree.tree[self.tree.P:self.tree.P + n] = prio
        self.tree.rebuild()

    def get_batch(self, idx, total="tree", dequant=None):
        idx = np.asarray(idx, np.int64)
        s, a, r, sp, done, _ = synthetic_transitions(self.seed, idx, tuple(self.spec["obs"]), self.spec["u8"], self.spec["nA"], self.alpha, self.eps)
        assert np.array_equal(a, self._a[idx]) and np.array_equal(r, self._r[idx])
        if dequant is not None:
            s, sp = dequant(s), dequant(sp)
        p = self._priorities[idx] / np.float32(self.tree.total)
        w = O.pow_f32(np.float32(self.n) * p, -self.beta)
        return s, a.astype(np.int64), r, sp, done.astype(np.float32), idx, w

    def update_priorities(self, idx, td):
        new = O.pow_f32(np.abs(np.asarray(td, np.float32)) + self.eps, self.alpha)
        assert np.all(new > 0)
        self._priorities[idx] = new
        self.tree.set_leaves(idx, new)


@pytest.fixture(scope="module")
def shard():
    return SyntheticShard(util.SPECS["c3_conv"], N, REPLAY_SEED)


@pytest.mark.parametrize("math_mode", [1, 0], ids=["3xtf32", "fp32"])
def test_bench_config_tree_sample_gather_and_step(lib, shard, math_mode):
    spec = dict(util.SPECS["c3_conv"]); spec["N"] = N
    cfg = lib.make_config(util.layer_descs(spec), (84, 84, 4), 6, obs_dtype="u8", batch_size=256, buffer_size=N, learning_rate=1e-4,
                          discount=0.99, seed=SEED, math_mode=math_mode)
    eng = lib.Engine(cfg)
    net = util.make_oracle_net(spec, True, seed=1)
    tgt = util.perturbed_copy(net, seed=22)
    eng.set_params(O.flat_params(net), 0)
    eng.set_params(O.flat_params(tgt), 1)
    eng.replay_fill_synthetic(N, REPLAY_SEED)
    assert eng.replay_size() == (N, 0)
    # priorities and the whole tree
    buf = SyntheticShard.__new__(SyntheticShard)          # private copy of the oracle state (the steps below update priorities)
    buf.__dict__.update(shard.__dict__)
    buf._priorities = shard._priorities.copy()
    buf.tree = O.SumTree(N); buf.tree.tree[...] = shard.tree.tree
    assert np.array_equal(eng.get_priorities(), buf._priorities)
    assert np.array_equal(eng.get_tree()[1:], buf.tree.tree[1:])
    # sampling calls
    for call in (0, 1, 2, 3, 999, 2**33 + 7):
        want, _ = buf.tree.sample(256, SEED, call)
        got = eng.sample_indices(call)
        assert np.array_equal(got, want), f"sampling call {call}"
        assert len(set(got.tolist())) == 256
    # get_batch out of the 56 GB store
    idx, _ = buf.tree.sample(256, SEED, 5)
    s, a, r, sp, d, _, w = eng.get_batch(idx)
    so, ao, ro, spo, do, _, wo = buf.get_batch(idx, dequant=util.dequant)
    assert np.array_equal(s, so) and np.array_equal(sp, spo)
    assert np.array_equal(a, ao) and np.array_equal(r, ro) and np.array_equal(d, do)
    assert util.relerr(w, wo) < 1e-6
    # rows far apart in the store (first, last, 2^31-byte boundaries of the byte offsets)
    probe = np.array([0, 1, 76_094, 76_095, 152_190, N // 2, N - 2, N - 1], np.int64)     # 76 095 * 28 224 B ~ 2^31
    ps, pa, pr, psp, pd = eng.replay_read(probe)
    qs, qa, qr, qsp, qd, _ = synthetic_transitions(REPLAY_SEED, probe, (4, 84, 84), True, 6)
    assert np.array_equal(ps, qs) and np.array_equal(psp, qsp) and np.array_equal(pa, qa) and np.array_equal(pr, qr) and np.array_equal(pd, qd)
    # full steps, sampled on the device
    opt = O.Adam(spec["lr"])
    for call in range(2):
        check_step(spec, net, tgt, buf, eng, opt, call, True, True)
    eng.close()
