"""GPU parity of the recurrent path: EpisodeReplayBuffer + LSTM batch_train! (src/solver.jl:239-287,
src/episode_replay.jl:21-95) against oracle/recurrent.py on identical episodes.

  * sampled episode indices and start offsets: bit-exact; trace mask, actions, rewards of the gathered batch: bit-exact
  * Q-values over all T x B rows, targets, td, loss: <= 1e-5 normwise (tanh / sigmoid networks: no sub-gradient flips)
  * gradients per parameter array: <= 5e-5; state0 (h0, c0) receives no update (oracle/recurrent.py docstring)
  * Adam on the engine's own gradient: <= 2 ulp; acting with a carried hidden state (dqn_q_values + dqn_policy_reset)
"""
import numpy as np
import pytest

import oracle as O
import util

pytestmark = pytest.mark.gpu
SEED = 2


def build(lib, d, H, nA, T, B, dueling, double_q, math_mode, cap=48, L=14, lr=1e-3, hidden_dense=0, n_eps=None, seed=5):
    rng = np.random.default_rng(seed)
    dense = ([(H, hidden_dense, O.ACT_TANH), (hidden_dense, nA, O.ACT_IDENTITY)] if hidden_dense else [(H, nA, O.ACT_IDENTITY)])
    net = O.make_recurrent_q(d, H, dense, dueling, rng)
    tgt = O.make_recurrent_q(d, H, dense, dueling, rng)
    for p in net.params() + tgt.params():
        if p.ndim == 1:
            p += rng.normal(0, 0.05, p.shape).astype(np.float32)
    layers = [dict(kind=2, act=0, in_=0, out=0), dict(kind=3, act=0, in_=d, out=H)] + [dict(kind=0, act=a, in_=i, out=o) for i, o, a in dense]
    cfg = lib.make_config(layers, (d,), nA, obs_dtype="f32", dueling=dueling, double_q=double_q, prioritized_replay=False, batch_size=B,
                          buffer_size=cap, learning_rate=lr, discount=0.95, seed=SEED, math_mode=math_mode, trace_length=T, max_episode_length=L)
    eng = lib.Engine(cfg)
    assert eng.num_params == sum(p.size for p in net.params())
    eng.set_params(np.concatenate([p.ravel() for p in net.params()]), 0)
    eng.set_params(np.concatenate([p.ravel() for p in tgt.params()]), 1)
    buf = O.EpisodeReplayBuffer((d,), cap, B, T, L)
    for _ in range(n_eps or cap + 5):                                        # wraps the ring
        n = int(rng.integers(1, L + 1))
        s = rng.normal(size=(n, d)).astype(np.float32); sp = rng.normal(size=(n, d)).astype(np.float32)
        a = rng.integers(1, nA + 1, n).astype(np.int32); r = rng.uniform(-1, 1, n).astype(np.float32)
        done = np.zeros(n, np.uint8); done[-1] = 1
        eng.episode_add(s, a, r, sp, done)
        buf.add_episode(s, a, r, sp, done)
    return net, tgt, buf, eng


def check_recurrent_step(lib, net, tgt, buf, eng, call, double_q, lr, qtol=1e-5, gtol=5e-5):
    T, B = buf.trace_length, buf.batch_size
    theta0 = eng.get_params(0)
    m0, v0, bp0 = eng.get_adam_state()
    o = 0
    for p in net.params():
        p[...] = theta0[o:o + p.size].reshape(p.shape); o += p.size
    idx, start = buf.sample_indices(SEED, call)
    s, a, r, sp, done, mask = buf.get_batch(idx, start)
    loss, gn = eng.train_step()
    assert np.array_equal(eng.last_indices()[:B], idx)
    assert np.array_equal(eng.is_weights().reshape(T, B), mask.astype(np.float32))
    out = O.forward_backward_recurrent(net, tgt, s, a - 1, r, sp, done, mask, 0.95, double_q, np.float32)
    out64 = O.forward_backward_recurrent(net, tgt, s, a - 1, r, sp, done, mask, 0.95, double_q, np.float64)
    nA = out["q"].shape[-1]
    scale = max(np.abs(out["q"]).max(), np.abs(out["q_target_sp"]).max())
    for got, key in ((eng.q(0), "q"), (eng.q(1), "q_online_sp"), (eng.q(2), "q_target_sp")):
        assert np.abs(got.reshape(T, B, nA) - out[key]).max() <= qtol * scale, key
    y, best = eng.targets()
    same = (best.reshape(T, B) - 1) == out["best_a"]
    assert same.mean() > 0.97
    assert np.abs(y.reshape(T, B)[same] - out["y"][same]).max() <= qtol * scale
    td = eng.td().reshape(T, B)
    live = same & (mask == 1)
    assert np.abs(td[live] - out["td"][live]).max() <= 2 * qtol * scale
    if same[mask == 1].all():
        assert abs(loss - out["loss"]) <= 1e-5 * max(abs(out["loss"]), 1e-3)
        g = eng.grads()
        o = 0
        gmax = 0.0
        for k, (g32, g64) in enumerate(zip(out["grads"], out64["grads"])):
            sl = slice(o, o + g32.size); o += g32.size
            if k in (3, 4):                                                  # state0: no gradient reaches it in the reference
                assert np.all(g[sl] == 0)
                continue
            den = max(np.abs(g64).max(), 1e-6 * max(np.abs(x).max() for x in out64["grads"]))
            assert np.abs(g[sl] - g64.ravel()).max() <= gtol * den, ("grad array", k)
            gmax = max(gmax, float(np.abs(g64).max()))
        assert abs(gn - gmax) <= gtol * gmax and gn == np.float32(np.abs(g).max())
    # Adam on the engine's own gradient
    g = eng.grads()
    ps = [p.copy() for p in net.params()]
    adam = O.Adam(lr)
    gl, o = [], 0
    for k, p in enumerate(ps):
        adam.state[k] = [m0[o:o + p.size].reshape(p.shape).copy(), v0[o:o + p.size].reshape(p.shape).copy(), [bp0[0], bp0[1]]]
        gl.append(g[o:o + p.size].reshape(p.shape)); o += p.size
    adam.apply(ps, gl)
    want = np.concatenate([p.ravel() for p in ps])
    ulp = np.spacing(np.abs(want).astype(np.float32)) + np.float32(1e-12)
    assert (np.abs(eng.get_params(0) - want) <= 2 * ulp).all()
    return loss


@pytest.mark.parametrize("dueling,double_q,math_mode", [(False, True, 0), (True, True, 0), (True, False, 0), (True, True, 1)])
def test_recurrent_step_parity_small(lib, dueling, double_q, math_mode):
    net, tgt, buf, eng = build(lib, 12, 16, 4, 6, 5, dueling, double_q, math_mode)
    assert eng.episode_count() == (buf._curr_size, buf._idx)
    for call in (0, 1, 7, 2**33 + 3):
        idx, start = eng.episode_sample(call)
        widx, wstart = buf.sample_indices(SEED, call)
        assert np.array_equal(idx, widx) and np.array_equal(start, wstart), call
    for call in range(3):
        check_recurrent_step(lib, net, tgt, buf, eng, call, double_q, 1e-3)
        if call == 0:
            eng.sync_target()
            o = 0
            th = eng.get_params(0)
            for p in tgt.params():
                p[...] = th[o:o + p.size].reshape(p.shape); o += p.size
    eng.close()


@pytest.mark.parametrize("math_mode", [0, 1], ids=["fp32", "3xtf32"])
def test_recurrent_step_parity_config4(lib, math_mode):
    # BASELINE.json configs[3]: LSTM-128 head on dim-128 observations, |A| = 16, trace length 32, batch 64 (episodes of <= 100 steps)
    net, tgt, buf, eng = build(lib, 128, 128, 16, 32, 64, True, True, math_mode, cap=96, L=100, lr=1e-4, n_eps=96, seed=9)
    for call in range(2):
        check_recurrent_step(lib, net, tgt, buf, eng, call, True, 1e-4)
    eng.close()


@pytest.mark.parametrize("H,B,math_mode", [(64, 5, 0), (256, 11, 0), (64, 9, 1)])
def test_recurrent_step_parity_cluster_kernels(lib, H, B, math_mode):
    """hidden sizes the one-launch recurrence takes (lstm_seq_*_kernel: clusters of 8 CTAs, 8 batch rows each): partial last cluster,
    several clusters, both contraction modes around it"""
    net, tgt, buf, eng = build(lib, 20, H, 5, 7, B, True, True, math_mode, cap=40, L=12, lr=1e-3, seed=13)
    for call in range(3):
        check_recurrent_step(lib, net, tgt, buf, eng, call, True, 1e-3)
    eng.close()


def test_recurrent_per_step_kernels_at_config4(lib, monkeypatch):
    """the fallback (one launch per time step: any hidden size) at the size where the cluster kernels normally run"""
    monkeypatch.setenv("DQN_LSTM_SEQ", "0")
    net, tgt, buf, eng = build(lib, 128, 128, 16, 32, 64, True, True, 0, cap=96, L=100, lr=1e-4, n_eps=96, seed=9)
    check_recurrent_step(lib, net, tgt, buf, eng, 0, True, 1e-4)
    eng.close()


def test_recurrent_acting_carries_hidden_state(lib):
    net, tgt, buf, eng = build(lib, 12, 16, 4, 6, 5, True, True, 0)
    rng = np.random.default_rng(3)
    for episode in range(2):
        eng.policy_reset(); net.reset()
        for t in range(5):
            obs = rng.normal(size=(3, 12)).astype(np.float32)               # three lanes, each with its own hidden state
            q = eng.q_values(obs)
            want = net(obs)
            assert np.abs(q - want).max() <= 1e-5 * np.abs(want).max(), (episode, t)
    with pytest.raises(lib.DQNError):
        eng.replay_add(np.zeros((1, 12), np.float32), [1], [0.0], np.zeros((1, 12), np.float32), [0], [0.1])
    eng.close()


def test_recurrent_errors(lib):
    net, tgt, buf, eng = build(lib, 12, 16, 4, 6, 5, True, True, 0, n_eps=3)
    with pytest.raises(lib.DQNError) as ei:                                  # episode_replay.jl:73 @assert r._curr_size >= r.batch_size
        eng.train_step()
    assert ei.value.code == lib._capi.DQN_ERR_STATE
    with pytest.raises(lib.DQNError):
        eng.episode_add(np.zeros((20, 12), np.float32), np.ones(20, np.int32), np.zeros(20, np.float32), np.zeros((20, 12), np.float32), np.zeros(20, np.uint8))
    eng.close()
