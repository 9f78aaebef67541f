"""Vectorised acting on the device: dqn_act / dqn_act_device = forward + dueling combine + first-max argmax +
epsilon-greedy draw (src/solver.jl:83, src/policy.jl:38-46), and the device-resident lane loop act -> synthetic env step ->
dqn_replay_add_device that BASELINE.json configs[4] describes."""
import ctypes as C

import numpy as np
import pytest

import oracle as O
import util
from oracle.philox import sample_uniforms
from test_gpu_parity import setup_pair, make_engine

pytestmark = pytest.mark.gpu
ACT_KEY = 0xAC7105EED


def expected_actions(q, seed, call, eps, lane0=0):
    n, nA = q.shape
    lanes = np.arange(lane0, lane0 + n, dtype=np.uint32)
    u0 = sample_uniforms(seed ^ ACT_KEY, call, lanes, np.zeros(n, np.uint32))
    u1 = sample_uniforms(seed ^ ACT_KEY, call, lanes, np.ones(n, np.uint32))
    greedy = np.argmax(q, axis=1)
    rnd = np.minimum((u1 * np.float32(nA)).astype(np.float32).astype(np.int64), nA - 1)
    explore = (u0 < np.float32(eps)) if eps > 0 else np.zeros(n, bool)
    return np.where(explore, rnd, greedy) + 1, explore


@pytest.mark.parametrize("name,math_mode", [("c1_gridworld", 0), ("conv_small", 0), ("conv_tanh", 1), ("c3_conv", 1)])
def test_act_matches_argmax_and_exploration_stream(lib, name, math_mode):
    kw = dict(math_mode=math_mode)
    if name == "c3_conv":
        kw.update(n_fill=300, B=64)
    spec, net, tgt, buf, eng = setup_pair(lib, name, **kw)
    s, _, _, _, _ = util.random_transitions(spec, 150, seed=77)             # more lanes than one forward chunk for the small specs
    want_q = net(util.dequant(s))
    a0, q = eng.act(s, eps=0.0, call=3, want_q=True)
    assert np.abs(q - want_q).max() <= 1e-5 * np.abs(want_q).max()
    top2 = np.sort(want_q, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-4 * np.abs(want_q).max()
    assert np.array_equal(a0[clear] - 1, np.argmax(want_q, axis=1)[clear]) and clear.mean() > 0.9
    assert np.array_equal(a0 - 1, np.argmax(q, axis=1))                     # first maximal index of the engine's own Q
    for eps, call in ((0.3, 0), (1.0, 2**33 + 1)):
        a = eng.act(s, eps=eps, call=call)
        want, explore = expected_actions(q, 2, call, eps)
        assert np.array_equal(a, want)
        assert eps < 1.0 or explore.all()
    eng.close()


def test_device_lanes_feed_the_replay_without_the_host(lib):
    """512 synthetic lanes: act on device observations -> env step on the device -> dqn_replay_add_device; the ring then holds exactly
    the transitions the lanes produced (read back through dqn_replay_read)."""
    import torch
    spec = dict(util.SPECS["c3_conv"]); spec["B"] = 32; spec["N"] = 2048
    eng = make_engine(lib, spec, B=32, N=2048, math_mode=1)
    net = util.make_oracle_net(spec, True, seed=21)
    eng.set_params(O.flat_params(net), 0); eng.sync_target()
    lanes, elems = 512, 84 * 84 * 4
    dev = torch.device("cuda")
    obs = torch.randint(0, 256, (lanes, elems), dtype=torch.uint8, device=dev)          # engine layout (H, W, C) per lane
    nxt = torch.empty_like(obs)
    act = torch.empty(lanes, dtype=torch.int32, device=dev)
    rew = torch.empty(lanes, dtype=torch.float32, device=dev); td0 = torch.empty_like(rew)
    done = torch.empty(lanes, dtype=torch.uint8, device=dev)
    L = lib._capi.lib
    p = lambda t: C.c_void_p(t.data_ptr())
    torch.cuda.synchronize()
    kept = []
    for step in range(3):
        assert L.dqn_act_device(eng.h, p(obs), lanes, 1, 0.1, step, p(act), None) == 0
        assert L.dqn_synth_env_step(eng.h, p(nxt), p(rew), p(done), p(td0), lanes, 99, step) == 0
        # the store keeps s / s' in the engine layout: ingest expects Flux layout (C,H,W) -> hand it the transposed views
        s_flux = obs.view(lanes, 84, 84, 4).permute(0, 3, 1, 2).contiguous(); sp_flux = nxt.view(lanes, 84, 84, 4).permute(0, 3, 1, 2).contiguous()
        torch.cuda.synchronize()
        assert L.dqn_replay_add_device(eng.h, p(s_flux), p(act), p(rew), p(sp_flux), p(done), p(td0), lanes) == 0
        torch.cuda.synchronize()
        kept.append((s_flux.cpu().numpy(), act.cpu().numpy(), rew.cpu().numpy(), sp_flux.cpu().numpy(), done.cpu().numpy()))
        obs, nxt = nxt, obs
    assert eng.replay_size() == (3 * lanes, 3 * lanes)
    for step, (s, a, r, sp, d) in enumerate(kept):
        idx = np.arange(step * lanes, (step + 1) * lanes, 37)
        rs, ra, rr, rsp, rd = eng.replay_read(idx)
        j = idx - step * lanes
        assert np.array_equal(rs, s[j].reshape(rs.shape)) and np.array_equal(rsp, sp[j].reshape(rs.shape))
        assert np.array_equal(ra, a[j]) and np.array_equal(rr, r[j]) and np.array_equal(rd, d[j])
        assert ra.min() >= 1 and ra.max() <= 6
    # greedy part of the lanes' actions = argmax of the engine's own Q on the same observations
    a_host, q = eng.act(kept[0][0], eps=0.0, call=0, want_q=True)
    want, explore = expected_actions(q, 2, 0, 0.1)
    assert np.array_equal(kept[0][1], want)
    loss, gn = eng.train_step()
    assert np.isfinite(loss)
    eng.close()
