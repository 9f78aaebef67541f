"""GPU parity tests: the CUDA path, driven through the C-ABI, against the oracle on identical seeded inputs.

Tolerances (stated once, used everywhere):
  * sampled leaf indices, ring contents, actions, best_a, sum-tree nodes: BIT-EXACT
  * Q-values, TD errors, targets, IS weights, loss: normwise relative error <= 1e-5  (north_star)
  * gradients, per parameter array, max-norm and L2: <= 5e-5 relative on EVERY network, ReLU networks included.  A ReLU unit
    whose pre-activation lies within rounding distance of zero flips its sub-gradient between ANY two fp32 evaluations (a
    different summation order is enough - Flux against itself under another BLAS threading would do it;
    scripts/relu_flip_study.py shows it on the CPU alone), so the oracle's reverse pass takes the ReLU masks from the engine's
    own forward activations (dqn_get_activation -> mask_override, oracle/nets.py) and everything else is held to 5e-5: a 1e-3
    error in any weight- or input-gradient kernel fails the test.  The forward values themselves are compared without any
    such help (Q within 1e-5, normwise: max|dQ| / max|Q|)
  * Adam: given the engine's own gradients, the updated parameters match the oracle's Flux-Adam to 1 ulp
"""
import glob
import os

import numpy as np
import pytest

import oracle as O
import util

pytestmark = pytest.mark.gpu

SEED = 2
QTOL = 1e-5
GTOL = 5e-5


def make_engine(lib, spec, dueling=True, double_q=True, per=True, use_graph=True, math_mode=0, B=None, N=None, gamma=0.99):
    cfg = lib.make_config(util.layer_descs(spec), tuple(reversed(spec["obs"])), spec["nA"], obs_dtype="u8" if spec["u8"] else "f32",
                          dueling=dueling, double_q=double_q, prioritized_replay=per, batch_size=B or spec["B"],
                          buffer_size=N or spec["N"], learning_rate=spec["lr"], discount=gamma, seed=SEED, use_graph=use_graph,
                          math_mode=math_mode)
    return lib.Engine(cfg)


def setup_pair(lib, name, dueling=True, double_q=True, per=True, n_fill=None, **kw):
    spec = dict(util.SPECS[name])
    if "B" in kw and kw["B"]:
        spec["B"] = kw["B"]
    net = util.make_oracle_net(spec, dueling, seed=21)
    tgt = util.perturbed_copy(net, seed=22)
    buf = util.make_oracle_replay(spec)
    eng = make_engine(lib, spec, dueling, double_q, per, **kw)
    assert eng.num_params == O.num_params(net)
    eng.set_params(O.flat_params(net), 0)
    eng.set_params(O.flat_params(tgt), 1)
    n = n_fill or (spec["N"] + 37)                       # wraps the ring
    s, a, r, sp, done = util.random_transitions(spec, n, seed=23)
    td0 = np.abs(r)
    k = n // 3
    for lo, hi in ((0, k), (k, n)):                      # two calls: cursor carried across calls
        eng.replay_add(s[lo:hi], a[lo:hi], r[lo:hi], sp[lo:hi], done[lo:hi], td0[lo:hi])
    buf.add_batch(s, a, r, sp, done, td0)
    return spec, net, tgt, buf, eng


def test_params_roundtrip_and_layout(lib):
    for name, dueling in (("c1_gridworld", True), ("conv_small", True), ("conv_small", False), ("c3_conv", True)):
        spec = util.SPECS[name]
        eng = make_engine(lib, spec, dueling, N=64, B=8)
        flat = np.random.default_rng(0).normal(size=eng.num_params).astype(np.float32)
        eng.set_params(flat, 0)
        assert np.array_equal(eng.get_params(0), flat)
        eng.sync_target()
        assert np.array_equal(eng.get_params(1), flat)
        eng.close()


@pytest.mark.parametrize("name", ["c1_gridworld", "testmdp", "conv_small"])
def test_replay_ring_priorities_and_tree_bit_exact(lib, name):
    spec, net, tgt, buf, eng = setup_pair(lib, name)
    assert eng.replay_size() == (buf._curr_size, buf._idx)
    assert np.array_equal(eng.get_priorities(), buf._priorities[:buf._curr_size])
    tree = eng.get_tree()
    assert np.array_equal(tree[1:], buf.tree.tree[1:])                      # every node, bit for bit
    idx = np.arange(0, buf._curr_size, max(1, buf._curr_size // 50))
    s, a, r, sp, d = eng.replay_read(idx)
    assert np.array_equal(s, buf._s[idx]) and np.array_equal(sp, buf._sp[idx])
    assert np.array_equal(a, buf._a[idx]) and np.array_equal(r, buf._r[idx]) and np.array_equal(d, buf._done[idx])
    # update_priorities! (PER.jl:76-80)
    td = np.random.default_rng(3).normal(0, 1, 20).astype(np.float32)
    ii = np.random.default_rng(4).choice(buf._curr_size, 20, replace=False)
    eng.update_priorities(ii, td)
    buf.update_priorities(ii, td)
    assert np.array_equal(eng.get_tree()[1:], buf.tree.tree[1:])
    eng.close()


@pytest.mark.parametrize("name", ["c1_gridworld", "testmdp", "conv_small", "c2_mlp"])
def test_sampled_indices_bit_exact(lib, name):
    spec, net, tgt, buf, eng = setup_pair(lib, name)
    for call in (0, 1, 2, 17, 2**33 + 5):
        got = eng.sample_indices(call)
        want, _ = buf.tree.sample(spec["B"], SEED, call)
        assert np.array_equal(got, want), f"call {call}"
        assert len(set(got.tolist())) == spec["B"]
    eng.close()


def test_sampling_small_buffer_forces_redraws(lib):
    # README-scale: n barely above B => many duplicates in round 0 (SURVEY 7 'Without-replacement semantics')
    spec, net, tgt, buf, eng = setup_pair(lib, "c1_gridworld", n_fill=40)
    for call in range(6):
        got = eng.sample_indices(call)
        want, att = buf.tree.sample(32, SEED, call)
        assert np.array_equal(got, want)
    assert att.max() >= 1
    eng.close()


@pytest.mark.parametrize("name", ["c1_gridworld", "testmdp", "conv_small"])
def test_get_batch_matches_reference_get_batch(lib, name):
    spec, net, tgt, buf, eng = setup_pair(lib, name)
    idx, _ = buf.tree.sample(spec["B"], SEED, 5)
    s, a, r, sp, d, _, w = eng.get_batch(idx)
    so, ao, ro, spo, do, _, wo = buf.get_batch(idx, total="tree", dequant=util.dequant)
    assert np.array_equal(s, so) and np.array_equal(sp, spo)                # k/255f0 exactly
    assert np.array_equal(a, ao) and np.array_equal(r, ro) and np.array_equal(d, do)
    assert np.array_equal(w, wo) or util.relerr(w, wo) < 1e-6
    _, _, _, _, _, _, wj = buf.get_batch(idx, total="pairwise", dequant=util.dequant)   # Julia's pairwise sum(prio)
    assert util.relerr(w, wj) < 1e-6
    eng.close()


def set_relu_masks(spec, net, eng):
    """mask_override on every ReLU layer of the oracle network <- (engine activation > 0) on the s rows of the step just run"""
    shape = tuple(spec["obs"])
    convs = [l for l in spec["layers"] if l[0] == "conv"]
    nc = len(convs)
    dueling = isinstance(net, O.DuelingNetwork)
    all_layers = (net.base.layers if dueling else net.layers)
    stage = 0
    for l in all_layers:
        if isinstance(l, O.Conv):
            c, h, w = shape
            shape = (l.cout,) + l.out_hw(h, w)
            if l.act == O.ACT_RELU:
                l.mask_override = eng.activation(stage, 0, shape) > 0
            stage += 1
    towers = [net.val.layers, net.adv.layers] if dueling else [[l for l in net.layers if isinstance(l, O.Dense)]]
    for t, layers in enumerate(towers):
        for li, l in enumerate(layers):
            if l.act == O.ACT_RELU:
                l.mask_override = eng.activation(nc + li, t, (l.nout,)) > 0


def clear_relu_masks(net):
    layers = (net.base.layers + net.val.layers + net.adv.layers) if isinstance(net, O.DuelingNetwork) else net.layers
    for l in layers:
        if hasattr(l, "mask_override"):
            del l.mask_override


def check_step(spec, net, tgt, buf, eng, opt, call, double_q, per, gamma=0.99, qtol=QTOL, gtol=GTOL):
    gmax_tol = gtol
    B = spec["B"]
    theta_before = eng.get_params(0)
    m0, v0, bp0 = eng.get_adam_state()
    O.set_params(net, theta_before)                                          # teacher forcing (see module docstring)
    want_idx, _ = buf.tree.sample(B, SEED, call)
    loss, gn = eng.train_step()
    idx = eng.last_indices()
    assert np.array_equal(idx, want_idx)
    sb, ab, rb, spb, db, _, w = buf.get_batch(idx, total="tree", dequant=util.dequant)
    set_relu_masks(spec, net, eng)                                           # reverse pass through the engine's own ReLU masks
    out = O.forward_backward(net, tgt, sb, ab - 1, rb, spb, db, w, gamma, double_q, np.float32)
    out64 = O.forward_backward(net, tgt, sb, ab - 1, rb, spb, db, w, gamma, double_q, np.float64)
    clear_relu_masks(net)
    assert util.relerr(eng.is_weights(), w) < 1e-6
    q, qo, qt = eng.q(0), eng.q(1), eng.q(2)
    scale = max(np.abs(out["q"]).max(), np.abs(out["q_target_sp"]).max())
    for got, key in ((q, "q"), (qo, "q_online_sp"), (qt, "q_target_sp")):
        assert np.abs(got - out[key]).max() <= qtol * scale, key
    y, best = eng.targets()
    src = out["q_online_sp"] if double_q else out["q_target_sp"]
    top2 = np.sort(src, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4 * scale                         # argmax well separated
    assert np.array_equal(best[clear] - 1, out["best_a"][clear]) and clear.mean() > 0.9
    same = (best - 1) == out["best_a"]
    assert np.abs(y[same] - out["y"][same]).max() <= qtol * scale
    td = eng.td()
    assert np.abs(td[same] - out["td"][same]).max() <= 2 * qtol * scale
    if same.all():
        assert abs(loss - out["loss"]) <= 1e-5 * max(abs(out["loss"]), 1e-3)
        g = eng.grads()
        gref = np.concatenate([x.ravel() for x in out["grads"]])
        g64 = np.concatenate([x.ravel() for x in out64["grads"]])
        o = 0
        for arr in out["grads"]:
            sl = slice(o, o + arr.size)
            den = max(np.abs(g64[sl]).max(), 1e-6 * np.abs(g64).max())
            assert np.abs(g[sl] - gref[sl]).max() <= gmax_tol * den, ("grad array at", o)
            assert np.abs(g[sl] - g64[sl]).max() <= gmax_tol * den
            l2 = max(np.linalg.norm(g64[sl]), 1e-6 * np.linalg.norm(g64))
            assert np.linalg.norm(g[sl] - g64[sl]) <= gmax_tol * l2, ("grad array (L2) at", o)
            o += arr.size
        assert abs(gn - out["grad_norm"]) <= gmax_tol * out["grad_norm"]
        assert gn == np.float32(np.abs(g).max())                             # globalnorm is max|g| (helpers.jl:38-46)
    # Adam in isolation: oracle Flux-Adam applied to the engine's own gradient
    g = eng.grads()
    ps = [p.copy() for p in net.params()]
    shapes = [p.shape for p in ps]
    gl, o = [], 0
    for shp in shapes:
        n = int(np.prod(shp)); gl.append(g[o:o + n].reshape(shp)); o += n
    adam = O.Adam(spec["lr"])
    o = 0
    for k, shp in enumerate(shapes):
        n = int(np.prod(shp))
        adam.state[k] = [m0[o:o + n].reshape(shp).copy(), v0[o:o + n].reshape(shp).copy(), [bp0[0], bp0[1]]]
        o += n
    adam.apply(ps, gl)
    theta_after = eng.get_params(0)
    want = np.concatenate([p.ravel() for p in ps])
    ulp = np.spacing(np.abs(want).astype(np.float32)) + np.float32(1e-12)
    assert (np.abs(theta_after - want) <= 2 * ulp).all(), "Adam update differs from Flux.Optimise.Adam by more than 1-2 ulp"
    m1, v1, bp1 = eng.get_adam_state()
    assert bp1 == (bp0[0] * 0.9, bp0[1] * 0.999)
    assert np.array_equal(m1, np.concatenate([adam.state[k][0].ravel() for k in range(len(shapes))]))
    # priorities: (|td|+eps)^alpha from the engine's td, tree bit-exact
    if per:
        buf.update_priorities(idx, td)
    assert np.array_equal(eng.get_tree()[1:], buf.tree.tree[1:])
    return loss, gn


CASES = [("mlp_tanh", True, True, True), ("conv_tanh", True, True, True), ("c1_gridworld", True, True, True), ("c1_gridworld", False, True, True), ("c1_gridworld", True, False, False),
         ("testmdp", True, True, True), ("conv_small", True, True, True), ("conv_small", False, False, True), ("c2_mlp", True, True, True)]


@pytest.mark.parametrize("name,dueling,double_q,per", CASES)
def test_batch_train_step_parity(lib, name, dueling, double_q, per):
    spec, net, tgt, buf, eng = setup_pair(lib, name, dueling, double_q, per)
    opt = O.Adam(spec["lr"])
    for call in range(4):
        check_step(spec, net, tgt, buf, eng, opt, call, double_q, per)
        if call == 1:                                                        # Flux.loadparams!(target_q, ...) solver.jl:142-145
            eng.sync_target()
            O.set_params(tgt, eng.get_params(0))
            assert np.array_equal(eng.get_params(1), eng.get_params(0))
    eng.close()


@pytest.mark.parametrize("math_mode", [0, 1], ids=["fp32", "3xtf32"])
def test_batch_train_step_parity_c3_full_batch(lib, math_mode):
    # BASELINE.json configs[2] network at its full batch (256); small buffer so the oracle finishes in seconds
    spec, net, tgt, buf, eng = setup_pair(lib, "c3_conv", n_fill=600, math_mode=math_mode)
    opt = O.Adam(spec["lr"])
    for call in range(2):
        check_step(spec, net, tgt, buf, eng, opt, call, True, True)
    eng.close()


@pytest.mark.parametrize("name", ["conv_small", "c2_mlp", "testmdp", "mlp_tanh", "conv_tanh"])
def test_batch_train_step_parity_tcgen05(lib, name):
    # the tensor-core (3xTF32) path takes every contraction it covers; the rest stays on the fp32 kernels
    spec, net, tgt, buf, eng = setup_pair(lib, name, math_mode=1)
    opt = O.Adam(spec["lr"])
    for call in range(3):
        check_step(spec, net, tgt, buf, eng, opt, call, True, True)
    eng.close()


def test_graph_and_eager_are_bit_identical(lib):
    res = []
    for use_graph in (True, False):
        spec, net, tgt, buf, eng = setup_pair(lib, "conv_small", use_graph=use_graph)
        out = [eng.train_step() for _ in range(3)]
        res.append((out, eng.get_params(0), eng.get_tree(), eng.td()))
        eng.close()
    assert res[0][0] == res[1][0]
    for a, b in zip(res[0][1:], res[1][1:]):
        assert np.array_equal(a, b)


def test_train_step_with_indices_equals_sampled_step(lib):
    spec, net, tgt, buf, eng = setup_pair(lib, "testmdp")
    spec2, net2, tgt2, buf2, eng2 = setup_pair(lib, "testmdp")
    idx = eng.sample_indices(0)
    a = eng.train_step()
    b = eng2.train_step_with_indices(idx)
    assert a == b and np.array_equal(eng.get_params(0), eng2.get_params(0))
    eng.close(); eng2.close()


@pytest.mark.parametrize("ingest_lane", ["1", "0"])
def test_one_step_ahead_loop_equals_the_serial_loop(lib, monkeypatch, ingest_lane):
    """add; step; read  vs  add_{k+1}; launch_{k+1}; read_k (dqn_step_result back=1, copies on their own stream): the same
    transitions enter the ring at the same point of the device order, so every (loss, grad_norm) and the final state are BIT-identical."""
    # ingest lane on: the transitions of step k+1 are written into the ring beside the reverse pass of step k (behind its tree epoch);
    # off: behind the whole step.  Same order of effects either way.
    spec, net, tgt, buf, eng = setup_pair(lib, "conv_small")
    monkeypatch.setenv("DQN_INGEST_LANE", ingest_lane)
    spec2, net2, tgt2, buf2, eng2 = setup_pair(lib, "conv_small")
    K = 12
    batches = [util.random_transitions(spec, 5, seed=100 + k) for k in range(K)]
    serial, ahead = [], []
    for s, a, r, sp, done in batches:
        eng.replay_add(s, a, r, sp, done, np.abs(r))
        serial.append(eng.train_step())
    for k, (s, a, r, sp, done) in enumerate(batches):
        eng2.replay_add(s, a, r, sp, done, np.abs(r))
        eng2.train_step_async()
        if k:
            ahead.append(eng2.step_result(1))
    ahead.append(eng2.step_result(0))
    assert serial == ahead
    assert eng2.step_result(1) == serial[-2]                # the slot of the step before the latest is still intact
    assert np.array_equal(eng.get_params(0), eng2.get_params(0))
    assert np.array_equal(eng.sample_indices(7), eng2.sample_indices(7))
    with pytest.raises(lib.DQNError):
        eng2.step_result(2)
    eng.close(); eng2.close()


def test_ingest_lane_under_a_mixed_call_sequence(lib, monkeypatch):
    """A few hundred calls in random order - adds of 1..9 transitions while a step is in flight, steps (async and serial), scalar reads
    one step late, priority updates, index draws, target syncs, reads of the ring - on an engine with the ingest lane and on one that
    ingests on the main stream: every returned value and the final state are BIT-identical (the lane only changes what overlaps)."""
    spec = util.SPECS["conv_small"]
    engs = []
    for lane in ("1", "0"):
        monkeypatch.setenv("DQN_INGEST_LANE", lane)
        engs.append(setup_pair(lib, "conv_small")[4])
    rng = np.random.default_rng(77)
    ops = rng.integers(0, 8, 400)
    outs = [[], []]
    for k, op in enumerate(ops):
        n = int(rng.integers(1, 10))
        s, a, r, sp, done = util.random_transitions(spec, n, seed=1000 + k)
        idx = rng.choice(spec["N"], 7, replace=False).astype(np.int64)      # (distinct: duplicate leaves in one update would race)
        td = rng.uniform(-2, 2, 7).astype(np.float32)
        for e, out in zip(engs, outs):
            if op <= 2:
                e.replay_add(s, a, r, sp, done, np.abs(r))
            elif op == 3:
                e.train_step_async()
            elif op == 4:
                out.append(e.train_step())
            elif op == 5:
                e.update_priorities(idx, td)
                out.append(tuple(e.sample_indices(k)))
            elif op == 6:
                e.sync_target() if k % 3 == 0 else out.append(tuple(e.replay_size()))
            else:
                e.train_step_async(); e.replay_add(s, a, r, sp, done, np.abs(r)); e.train_step_async(); out.append(e.step_result(1))
    for e, out in zip(engs, outs):
        out.append(e.sync())
    assert outs[0] == outs[1]
    assert np.array_equal(engs[0].get_params(0), engs[1].get_params(0)) and np.array_equal(engs[0].get_params(1), engs[1].get_params(1))
    assert np.array_equal(engs[0].get_tree(), engs[1].get_tree())
    for e in engs:
        e.close()


@pytest.mark.parametrize("name,dueling", [("c1_gridworld", True), ("conv_small", True), ("conv_small", False), ("testmdp", True)])
def test_acting_q_values(lib, name, dueling):
    spec, net, tgt, buf, eng = setup_pair(lib, name, dueling)
    s, a, r, sp, done = util.random_transitions(spec, 70, seed=31)         # more rows than one chunk for c1
    q = eng.q_values(s, 0)
    want = net(util.dequant(s))
    assert np.abs(q - want).max() <= QTOL * np.abs(want).max()
    qt = eng.q_values(s[:3], 1)
    assert np.abs(qt - tgt(util.dequant(s[:3]))).max() <= QTOL * np.abs(want).max()
    eng.close()


def test_error_codes_mirror_the_reference_asserts(lib):
    spec = util.SPECS["c1_gridworld"]
    eng = make_engine(lib, spec)
    s, a, r, sp, done = util.random_transitions(spec, 10, seed=1)
    eng.replay_add(s, a, r, sp, done, np.abs(r))
    with pytest.raises(lib.DQNError) as ei:                                  # PER.jl:83 @assert r._curr_size >= r.batch_size
        eng.train_step()
    assert ei.value.code == lib._capi.DQN_ERR_STATE
    with pytest.raises(lib.DQNError) as ei:                                  # PER.jl:66 @assert td_err + eps > 0
        eng.replay_add(s[:1], a[:1], r[:1], sp[:1], done[:1], np.array([-1.0], np.float32))
    assert ei.value.code == lib._capi.DQN_ERR_STATE
    with pytest.raises(lib.DQNError):
        eng.replay_add(s[:1], np.array([9], np.int32), r[:1], sp[:1], done[:1], np.abs(r[:1]))
    with pytest.raises(lib.DQNError):
        eng.set_params(np.zeros(3, np.float32))
    with pytest.raises(lib.DQNError):                                        # dueling on a chain without a trailing Dense (dueling.jl:47-50)
        lib.Engine(lib.make_config([dict(kind=2, act=0, in_=0, out=0)], (2,), 4))
    eng.close()


GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("ref_"))


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_engine_against_committed_golden_vectors(lib, path):
    g = np.load(path)
    base = os.path.basename(path)[:-4]
    name, flags = base.rsplit("_", 1)
    dueling, double_q = flags[1] == "1", flags[3] == "1"
    spec = util.SPECS[name]
    eng = make_engine(lib, spec, dueling, double_q)
    eng.set_params(g["theta0"], 0)
    eng.set_params(g["theta_t"], 1)
    eng.replay_add(g["s"], g["a"], g["r"], g["sp"], g["done"], np.abs(g["r"]))
    assert np.array_equal(eng.sample_indices(0), g["idx"])
    loss, gn = eng.train_step()
    assert np.array_equal(eng.last_indices(), g["idx"])
    scale = np.abs(g["q"]).max()
    assert np.abs(eng.q(0) - g["q"]).max() <= QTOL * scale
    assert np.abs(eng.q(0) - g["q64"]).max() <= QTOL * scale
    assert np.array_equal(eng.targets()[1] - 1, g["best_a"])
    assert np.abs(eng.td() - g["td"]).max() <= 2 * QTOL * scale
    assert abs(loss - g["loss"]) <= 1e-5 * abs(g["loss"])
    gt = 5e-3 if name == "conv_small" else GTOL                  # ReLU network: flip-aware bound (module docstring)
    assert util.relerr(eng.grads(), g["grads64"]) <= gt
    assert abs(gn - g["grad_norm"]) <= gt * g["grad_norm"]
    assert np.abs(eng.get_priorities() - g["prio1"][:eng.replay_size()[0]]).max() <= 1e-5
    refp = os.path.join(os.path.dirname(path), "ref_" + base + ".npz")       # outputs of the reference itself (oracle/julia/mint_fixtures.jl), when minted
    if os.path.exists(refp):
        r = np.load(refp)
        assert np.array_equal(eng.targets()[1] - 1, r["best_a"])
        assert np.abs(eng.q(0) - r["q"]).max() <= QTOL * scale and np.abs(eng.td() - r["td"]).max() <= 2 * QTOL * scale
        assert abs(loss - float(r["loss"][0])) <= 1e-5 * abs(float(r["loss"][0]))
        assert util.relerr(eng.grads(), r["grads"]) <= gt
    eng.close()


def test_learning_signal_fixed_batch(lib):
    # repeated steps on the same indices drive the loss down (sanity of sign conventions end to end)
    spec, net, tgt, buf, eng = setup_pair(lib, "testmdp")
    idx = eng.sample_indices(0)
    losses = [eng.train_step_with_indices(idx)[0] for _ in range(60)]
    assert losses[-1] < 0.5 * losses[0]
    eng.close()


def test_tc_selftest():
    """The tcgen05 kernel in isolation: every operand layout (K-major / MN-major A and B, parity-class dgrad, several tiles per
    persistent CTA) against the CPU executor of the same operand functor.  The binary is built by __graft_entry__.build()."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "tc_selftest")
    if not os.path.exists(exe):
        pytest.skip("tests/csrc/tc_selftest not built (run __graft_entry__.build())")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "SELFTEST OK" in r.stdout, r.stdout[-2000:] + r.stderr[-500:]


def test_env_switches_keep_parity(lib):
    """Schedule / kernel variants (merged online+target launches, single lane, tiled heads instead of the warp-per-row kernel, cp.async
    instead of the TMA feed, the generic kernel instead of the dedicated first-layer kernel, TMA-fed weight gradients) compute the same step."""
    import subprocess, sys
    code = ("import sys,os; sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'tests'); import numpy as np, dqn_b200 as lib, util, oracle as O;"
            "spec=util.SPECS['c3_conv']; net=util.make_oracle_net(spec, True, seed=21);"
            "cfg=lib.make_config(util.layer_descs(spec), tuple(reversed(spec['obs'])), spec['nA'], obs_dtype='u8', batch_size=64, buffer_size=512,"
            " learning_rate=1e-4, discount=0.99, seed=2, math_mode=1); e=lib.Engine(cfg); e.set_params(O.flat_params(net),0); e.sync_target();"
            "e.replay_fill_synthetic(512, 7); l,g=e.train_step(); print('RES', repr(float(l)), repr(float(g)), repr(float(np.abs(e.grads()).sum())))")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for tag, env in (("default", {}), ("merge", {"DQN_MERGE_FWD": "1"}), ("one_lane", {"DQN_STREAMS": "0"}), ("tiled_heads", {"DQN_FUSE_HEADS": "0"}), ("no_a8", {"DQN_NO_A8": "1"}),
                     ("tail_split", {"DQN_TC_TAIL": "1"}), ("cp_async_feed", {"DQN_TC_TMA": "0"}), ("generic_conv1", {"DQN_TC_C1": "0"}), ("tma_wgrad", {"DQN_TC_TMA_WGRAD": "1"}),
                     ("single_head_kernel", {"DQN_FUSE_HEAD_ALL": "1"}), ("head_loss_dgrad_kernel", {"DQN_FUSE_HEAD_ALL": "2"}), ("classwise_dgrad", {"DQN_DGRAD_MERGE": "0"}), ("cp_async_dgrad", {"DQN_TC_TMA_DGRAD": "0"}), ("generic_head", {"DQN_HEAD_SMALL": "0"}), ("gather_first", {"DQN_C1_DIRECT": "0"})):
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=root, env={**os.environ, **env})
        line = [x for x in r.stdout.splitlines() if x.startswith("RES")]
        assert line, (tag, r.stdout[-500:], r.stderr[-1500:])
        outs[tag] = [float(x) for x in line[0].split()[1:]]
    ref = outs["default"]
    assert outs["one_lane"] == ref, outs                                    # same kernels, same order of operations: bit-identical
    assert outs["gather_first"] == ref, outs                                # first layer from the gathered batch instead of the store: same bytes
    assert outs["generic_head"] == ref, outs                                # the register-resident head is the generic one, operation for operation
    for tag in ("merge", "tiled_heads", "tail_split", "no_a8", "cp_async_feed", "generic_conv1", "tma_wgrad", "single_head_kernel", "head_loss_dgrad_kernel", "classwise_dgrad", "cp_async_dgrad"):   # different summation order in a few contractions
        for a, b in zip(outs[tag], ref):
            assert abs(a - b) <= 2e-5 * abs(b), (tag, outs)


def test_sampling_curr_size_equals_batch_size_is_a_permutation(lib):
    # curr_size == batch_size: the reference's sample(...; replace=false) returns every index (PER:83-85); rejection alone does not
    # terminate here - the exact exclusion descent finishes the batch, bit-identical to the oracle's restatement of it
    spec, net, tgt, buf, eng = setup_pair(lib, "c1_gridworld", n_fill=32)
    used_exact = 0
    for call in range(12):
        got = eng.sample_indices(call)
        want, att = buf.tree.sample(32, SEED, call)
        assert np.array_equal(got, want), f"call {call}"
        assert sorted(got.tolist()) == list(range(32))
        used_exact += int((att >= 0x80000000).any())
    assert used_exact > 0
    loss, gn = eng.train_step()                                              # and the step runs on such a batch
    assert np.isfinite(loss) and sorted(eng.last_indices().tolist()) == list(range(32))
    eng.close()


def test_sampler_distribution_matches_successive_sampling(lib):
    """Inclusion frequencies of the device sampler over 30 000 sampling calls against the exact inclusion probabilities of successive
    sampling without replacement (what StatsBase's A-ExpJ realises, PER:85; tests/test_oracle_cpu.py checks the A-ExpJ restatement
    against the same enumeration)."""
    import itertools
    n, B = 10, 4
    spec = dict(util.SPECS["c1_gridworld"]); spec["B"] = B; spec["N"] = n
    eng = make_engine(lib, spec, B=B, N=n)
    s, a, r, sp, done = util.random_transitions(spec, n, seed=3)
    eng.replay_add(s, a, r, sp, done, np.abs(r))
    p = eng.get_priorities().astype(np.float64)
    incl = np.zeros(n)
    for seq in itertools.permutations(range(n), B):
        rem, pr = p.sum(), 1.0
        for i in seq:
            pr *= p[i] / rem
            rem -= p[i]
        incl[list(seq)] += pr
    calls = 30000
    cnt = np.zeros(n)
    for call in range(calls):
        cnt[eng.sample_indices(call)] += 1
    f = cnt / calls
    sigma = np.sqrt(incl * (1 - incl) / calls)
    assert np.all(np.abs(f - incl) <= 4.5 * sigma + 1e-4), (f, incl)
    eng.close()


def test_replay_add_device_matches_host_add(lib):
    """dqn_replay_add_device (transitions already in HBM, the vectorised-env ingest path) fills the ring exactly like dqn_replay_add,
    and an offending transition (bad action / td0 + eps <= 0) is reported and not stored."""
    import ctypes as C
    import torch
    spec = util.SPECS["conv_small"]
    e_host = make_engine(lib, spec)
    e_dev = make_engine(lib, spec)
    n = spec["N"] + 11
    s, a, r, sp, done = util.random_transitions(spec, n, seed=41)
    td0 = np.abs(r)
    for lo, hi in ((0, 100), (100, n)):
        e_host.replay_add(s[lo:hi], a[lo:hi], r[lo:hi], sp[lo:hi], done[lo:hi], td0[lo:hi])
    dev = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    for lo, hi in ((0, 100), (100, n)):
        ts, ta, tr, tsp, td, ttd = dev(s[lo:hi]), dev(a[lo:hi]), dev(r[lo:hi]), dev(sp[lo:hi]), dev(done[lo:hi]), dev(td0[lo:hi])
        torch.cuda.synchronize()
        rc = lib._capi.lib.dqn_replay_add_device(e_dev.h, C.c_void_p(ts.data_ptr()), C.c_void_p(ta.data_ptr()), C.c_void_p(tr.data_ptr()),
                                                 C.c_void_p(tsp.data_ptr()), C.c_void_p(td.data_ptr()), C.c_void_p(ttd.data_ptr()), hi - lo)
        assert rc == 0, lib._capi.lib.dqn_last_error(e_dev.h)
    assert e_dev.replay_size() == e_host.replay_size()
    assert np.array_equal(e_dev.get_tree(), e_host.get_tree())
    idx = np.arange(0, spec["N"], 7)
    for x, y in zip(e_dev.replay_read(idx), e_host.replay_read(idx)):
        assert np.array_equal(x, y)
    assert np.array_equal(e_dev.sample_indices(3), e_host.sample_indices(3))
    # an invalid transition: flagged, skipped, everything else untouched
    before = e_dev.get_tree().copy()
    bad_a = dev(np.array([1, 99], np.int32)); ok_td = dev(np.array([0.5, 0.5], np.float32))
    ts, tr, tsp, td = dev(s[:2]), dev(r[:2]), dev(sp[:2]), dev(done[:2])
    torch.cuda.synchronize()
    rc = lib._capi.lib.dqn_replay_add_device(e_dev.h, C.c_void_p(ts.data_ptr()), C.c_void_p(bad_a.data_ptr()), C.c_void_p(tr.data_ptr()),
                                             C.c_void_p(tsp.data_ptr()), C.c_void_p(td.data_ptr()), C.c_void_p(ok_td.data_ptr()), 2)
    assert rc == lib._capi.DQN_ERR_INVALID
    after = e_dev.get_tree()
    P = after.size // 2
    cur = e_host.replay_size()[1]
    changed = np.nonzero(after[P:] != before[P:])[0]
    assert changed.tolist() in ([cur], [])                                   # only the valid transition's leaf (if its priority differs)
    loss, gn = e_dev.train_step()                                            # the sticky flag was consumed by the failed call
    assert np.isfinite(loss)
    e_host.close(); e_dev.close()


def test_data_parallel_two_gpus(n_gpus):
    """world = 2 under torchrun: the all-reduced gradient equals the oracle's gradient of the combined batch, parameters stay
    identical across ranks (scripts/mgpu_check.py; its output is kept under gpurun_out/ for profiles/)."""
    if n_gpus < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(root, "scripts", "mgpu_check.py")], capture_output=True, text=True, timeout=900, cwd=root)
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "mgpu_check.log"), "w") as f:
        f.write(r.stdout + "\n--- stderr ---\n" + r.stderr[-4000:])
    assert r.returncode == 0 and r.stdout.count("MGPU OK") == 2, r.stdout[-3000:] + r.stderr[-2000:]


def _group_case(lib, ndev):
    """dqn_group_create(ndev): one step on explicit per-shard data against the oracle's gradient of the combined batch."""
    spec = util.SPECS["conv_small"]
    B = spec["B"]
    net = util.make_oracle_net(spec, True, seed=21)
    tgt = util.perturbed_copy(net, seed=22)
    cfg = lib.make_config(util.layer_descs(spec), tuple(reversed(spec["obs"])), spec["nA"], obs_dtype="u8", batch_size=B, buffer_size=spec["N"],
                          learning_rate=spec["lr"], discount=0.99, seed=SEED)
    grp = lib.Group(cfg, ndev)
    grp.set_params(O.flat_params(net), 0)
    grp.set_params(O.flat_params(tgt), 1)
    bufs = []
    for r, eng in enumerate(grp.engines):
        s, a, rr, sp, done = util.random_transitions(spec, 150, seed=60 + r)
        eng.replay_add(s, a, rr, sp, done, np.abs(rr))
        buf = util.make_oracle_replay(spec); buf.add_batch(s, a, rr, sp, done, np.abs(rr)); bufs.append(buf)
    loss, gn = grp.train_step()
    S, A, R, SP, D, W, losses = [], [], [], [], [], [], []
    for r, (eng, buf) in enumerate(zip(grp.engines, bufs)):
        idx = eng.last_indices()
        want, _ = buf.tree.sample(B, SEED + r, 0)                            # rank r samples with seed + r
        assert np.array_equal(idx, want)
        sb, ab, rb, spb, db, _, w = buf.get_batch(idx, total="tree", dequant=util.dequant)
        S.append(sb); A.append(ab); R.append(rb); SP.append(spb); D.append(db); W.append(w)
    out = O.forward_backward(net, tgt, np.concatenate(S), np.concatenate(A) - 1, np.concatenate(R), np.concatenate(SP), np.concatenate(D),
                             np.concatenate(W), 0.99, True, np.float64)
    gref = np.concatenate([x.ravel() for x in out["grads"]])
    for eng in grp.engines:
        assert util.relerr(eng.grads(), gref) < 2e-4
    assert abs(loss - out["loss"]) <= 1e-5 * abs(out["loss"])
    th = [eng.get_params(0) for eng in grp.engines]
    for t in th[1:]:
        assert np.array_equal(t, th[0])                                      # identical Adam on every rank
    grp.close()


def test_group_of_one_is_a_plain_engine(lib):
    _group_case(lib, 1)


def test_group_single_process_two_gpus(lib, n_gpus):
    if n_gpus < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    _group_case(lib, 2)
