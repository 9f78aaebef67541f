"""The recurrent restatement (oracle/recurrent.py) pinned on the CPU: BPTT gradients against torch autograd in fp64, the episode
sampler's quirk (src/episode_replay.jl:81-92), and the loss normalisation of src/solver.jl:279-281."""
import numpy as np
import pytest
import torch

import oracle as O
from oracle.recurrent import forward_backward_recurrent, make_recurrent_q


@pytest.mark.parametrize("dueling", [False, True])
def test_bptt_gradients_match_torch_autograd(dueling):
    rng = np.random.default_rng(0)
    T, B, d, H, nA = 5, 6, 7, 8, 4
    net = make_recurrent_q(d, H, [(H, nA, O.ACT_IDENTITY)], dueling, rng)
    tgt = make_recurrent_q(d, H, [(H, nA, O.ACT_IDENTITY)], dueling, rng)
    s = rng.normal(size=(T, B, d)).astype(np.float32); sp = rng.normal(size=(T, B, d)).astype(np.float32)
    a = rng.integers(0, nA, (T, B)); r = rng.normal(size=(T, B)).astype(np.float32)
    done = (rng.uniform(size=(T, B)) < 0.2).astype(np.float32); mask = (rng.uniform(size=(T, B)) < 0.7).astype(np.int32)
    out = forward_backward_recurrent(net, tgt, s, a, r, sp, done, mask, 0.95, True, np.float64)
    P = [torch.tensor(p.astype(np.float64), requires_grad=True) for p in net.params()]
    Wi, Wh, b, h0, c0 = P[:5]
    rest = P[5:]
    h, c, loss = h0.expand(B, H), c0.expand(B, H), 0
    for t in range(T):
        g = torch.tensor(s[t].astype(np.float64)) @ Wi + h @ Wh + b
        i, f, gc, o = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
        c = f * c + i * gc
        h = o * torch.tanh(c)
        if dueling:
            v = h @ rest[0] + rest[1]; A = h @ rest[2] + rest[3]; q = (v + A) - A.mean(dim=1, keepdim=True)
        else:
            q = h @ rest[0] + rest[1]
        td = q[torch.arange(B), torch.tensor(a[t])] - torch.tensor(out["y"][t])
        x = torch.tensor(mask[t].astype(np.float64)) * td
        ax = x.abs(); quad = torch.clamp(ax, max=1.0)
        loss = loss + (0.5 * quad * quad + (ax - quad)).sum() / B
    loss = loss / T
    loss.backward()
    assert abs(float(loss.detach()) - float(out["loss"])) < 1e-12
    for p, g in zip(P, out["grads"]):
        ref = p.grad.numpy()
        assert np.abs(ref - g).max() <= 1e-9 * max(np.abs(ref).max(), 1e-30)


def test_episode_sampler_start_offset_quirk():
    # ep_start only shortens the trace: the copied steps are always ep[1], ep[2], ... (src/episode_replay.jl:81-92)
    buf = O.EpisodeReplayBuffer((3,), 10, 4, 6)
    rng = np.random.default_rng(1)
    for e in range(7):
        n = 3 + e
        s = np.full((n, 3), e, np.float32) + np.arange(n)[:, None] / 100
        buf.add_episode(s, np.ones(n, np.int32), np.arange(n, dtype=np.float32), s + 1, np.r_[np.zeros(n - 1), 1])
    idx, start = buf.sample_indices(2, 0)
    assert len(set(idx.tolist())) == 4 and np.all(idx < 7) and np.all(start >= 1)
    s, a, r, sp, d, m = buf.get_batch(idx, start)
    for i, (e, st) in enumerate(zip(idx, start)):
        n = 3 + e
        cnt = max(min(n, 6) - st + 1, 0)
        assert m[:, i].sum() == cnt and np.all(m[:cnt, i] == 1)
        assert np.array_equal(r[:cnt, i], np.arange(cnt))               # from the first step of the episode, not from ep_start
        assert np.all(s[cnt:, i] == 0) and np.all(a[cnt:, i] == 1)
