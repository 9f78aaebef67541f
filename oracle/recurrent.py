"""Recurrent batch_train! restatement (TEST INFRASTRUCTURE): src/solver.jl:239-287, src/episode_replay.jl:21-95 of the reference.

    LSTM                 Flux 0.14 Recur(LSTMCell): g = Wi x + Wh h + b, gate order input | forget | cell | output,
                         c' = sigma(f) c + sigma(i) tanh(g_c), h' = sigma(o) tanh(c'); forget-gate bias initialised to 1; state0 = (h0, c0)
                         zeros (out, 1) broadcast over the batch; Flux.reset! puts state = state0
    EpisodeReplayBuffer  ring of whole episodes; sample = batch_size distinct episodes, then the start-offset quirk of :81-92:
                         ep_start = rand(1:len) only SHORTENS the trace - steps ep[1], ep[2], ... are copied for j = ep_start:min(len, T)
    batch_train!         targets for t = 1..T with the online and the target network both stepping through sp_batch (state carried),
                         reset, then loss = (1/T) sum_t sum_i huber(mask_ti * td_ti) / B with BPTT through all T steps; no priorities

Array convention as oracle/nets.py (numpy image of the Julia arrays): Wi (4out,in) <-> numpy (in,4out), Wh (4out,out) <-> (out,4out),
b (4out,), h0/c0 (out,1) <-> (1,out); Flux.params order of the cell: Wi, Wh, b, h0, c0.

Two upstream points restated from memory of Flux 0.14 / Zygote (no source offline - "parity unpinned", flagged in DESIGN.md):
  * `_fast` activations (tanh_fast / sigmoid_fast) differ from tanh / sigmoid by a few ulp: inside the fp32 tolerance of the tests;
  * state0 receives NO gradient in this call pattern (reset! happens outside the gradient closure and the state tuple is reached through a
    mutable field, so the implicit-Params bookkeeping never sees the two arrays): h0, c0 stay at their initial zeros.  The oracle computes
    their gradients (`dstate0`) for reference but `batch_train_recurrent` does not apply them.
"""
import numpy as np

from .nets import Dense, Flatten, Chain, DuelingNetwork, _act, ACT_IDENTITY
from .philox import sample_uniforms
from .step import huber_loss, globalnorm, q_targets_of
from .sumtree import SumTree


def _sigmoid(z):
    one = z.dtype.type(1)
    return one / (one + np.exp(-z))


class LSTM:
    """Flux.LSTM(in, out) = Recur(LSTMCell)."""

    def __init__(self, nin, nout):
        self.nin, self.nout = int(nin), int(nout)
        self.Wi = np.zeros((nin, 4 * nout), np.float32)
        self.Wh = np.zeros((nout, 4 * nout), np.float32)
        self.b = np.zeros(4 * nout, np.float32)
        self.b[nout:2 * nout] = 1.0                       # cell.b[gate(out, 2)] .= 1
        self.h0 = np.zeros((1, nout), np.float32)
        self.c0 = np.zeros((1, nout), np.float32)
        self.state = None

    def params(self):
        return [self.Wi, self.Wh, self.b, self.h0, self.c0]

    def reset(self):
        self.state = None                                  # state = state0, broadcast at the next call

    def step(self, x, dtype):
        H = self.nout
        if self.state is None:
            h = np.broadcast_to(self.h0.astype(dtype), (x.shape[0], H))
            c = np.broadcast_to(self.c0.astype(dtype), (x.shape[0], H))
        else:
            h, c = self.state
        g = x @ self.Wi.astype(dtype) + h @ self.Wh.astype(dtype) + self.b.astype(dtype)
        i, f, o = _sigmoid(g[:, :H]), _sigmoid(g[:, H:2 * H]), _sigmoid(g[:, 3 * H:])
        gc = np.tanh(g[:, 2 * H:3 * H])
        c2 = f * c + i * gc
        tc = np.tanh(c2)
        h2 = o * tc
        self.state = (h2, c2)
        return h2, (x, h, c, i, f, gc, o, tc)

    def backward_step(self, dh, dc_next, cache):
        """dh: gradient into h' (heads + the next step's recurrence); dc_next: gradient into c' from the next step."""
        x, h, c, i, f, gc, o, tc = cache
        H = self.nout
        one = dh.dtype.type(1)
        do = dh * tc
        dc = dc_next + dh * o * (one - tc * tc)
        dg = np.concatenate([dc * gc * i * (one - i), dc * c * f * (one - f), dc * i * (one - gc * gc), do * o * (one - o)], axis=1)
        dWi = x.T @ dg
        dWh = h.T @ dg
        db = dg.sum(axis=0)
        dx = dg @ self.Wi.astype(dh.dtype).T
        dh_prev = dg @ self.Wh.astype(dh.dtype).T
        dc_prev = dc * f
        return dx, dh_prev, dc_prev, [dWi, dWh, db]


class RecurrentQ:
    """Chain(flattenbatch?, LSTM, Dense...) or its dueling split (base = [.., LSTM], val / adv = the trailing Dense layers, src/dueling.jl:36-58).
    Layers before the LSTM must be parameter-free (flattenbatch)."""

    def __init__(self, lstm, val_layers, adv_layers):
        self.lstm = lstm
        self.val = Chain(*val_layers) if val_layers is not None else None
        self.adv = Chain(*adv_layers)

    @property
    def dueling(self):
        return self.val is not None

    def params(self):           # Flux.params order: base (the LSTM cell), val, adv
        return self.lstm.params() + (self.val.params() if self.dueling else []) + self.adv.params()

    def reset(self):
        self.lstm.reset()

    def _heads(self, h, dtype):
        if not self.dueling:
            q, ca = self.adv.forward(h, dtype)
            return q, (None, ca, q.shape[1])
        v, cv = self.val.forward(h, dtype)
        a, ca = self.adv.forward(h, dtype)
        m = a[:, 0].copy()
        for k in range(1, a.shape[1]):
            m = m + a[:, k]
        mean = m / dtype(a.shape[1])
        return (v + a) - mean[:, None], (cv, ca, a.shape[1])

    def step(self, x, dtype=np.float32):
        x = np.asarray(x, dtype).reshape(x.shape[0], -1)
        h, cl = self.lstm.step(x, dtype)
        q, ch = self._heads(h, dtype)
        return q, (cl, ch)

    def __call__(self, x, dtype=np.float32):
        return self.step(x, dtype)[0]

    def backward_seq(self, dqs, caches):
        """dqs[t]: dL/dQ_t; caches[t] from step().  Returns gradients in params() order (state0 gradients included, see module docstring)."""
        T = len(dqs)
        H = self.lstm.nout
        dt = dqs[0].dtype
        gl = [np.zeros_like(p, dt) for p in self.lstm.params()]
        gv = [np.zeros_like(p, dt) for p in self.val.params()] if self.dueling else []
        ga = [np.zeros_like(p, dt) for p in self.adv.params()]
        dh_next = np.zeros((dqs[0].shape[0], H), dt)
        dc_next = np.zeros_like(dh_next)
        for t in range(T - 1, -1, -1):
            cl, (cv, ca, na) = caches[t]
            dq = dqs[t]
            if self.dueling:
                dv = dq[:, 0].copy()
                for k in range(1, na):
                    dv = dv + dq[:, k]
                dv = dv[:, None]
                da = dq - dv / dt.type(na)
                dhv, g1 = self.val.backward(dv, cv)
                dha, g2 = self.adv.backward(da, ca)
                dh = dhv + dha
                for a_, b_ in zip(gv, g1):
                    a_ += b_
            else:
                dh, g2 = self.adv.backward(dq, ca)
            for a_, b_ in zip(ga, g2):
                a_ += b_
            _, dh_next, dc_next, g3 = self.lstm.backward_step(dh + dh_next, dc_next, cl)
            for a_, b_ in zip(gl[:3], g3):
                a_ += b_
        gl[3] = dh_next.sum(axis=0, keepdims=True)
        gl[4] = dc_next.sum(axis=0, keepdims=True)
        return gl + gv + ga


def make_recurrent_q(nin, hidden, dense_specs, dueling, rng):
    """dense_specs: [(in, out, act), ...] trailing Dense layers of the Chain.  Glorot-uniform weights (Flux default), LSTM biases as Flux."""
    lstm = LSTM(nin, hidden)
    lim = np.sqrt(6.0 / (nin + 4 * hidden)); lstm.Wi[...] = rng.uniform(-lim, lim, lstm.Wi.shape).astype(np.float32)
    lim = np.sqrt(6.0 / (hidden + 4 * hidden)); lstm.Wh[...] = rng.uniform(-lim, lim, lstm.Wh.shape).astype(np.float32)

    def dense(i, o, act):
        d = Dense(i, o, act)
        lim = np.sqrt(6.0 / (i + o)); d.weight[...] = rng.uniform(-lim, lim, d.weight.shape).astype(np.float32)
        return d
    adv = [dense(*sp) for sp in dense_specs]
    val = None
    if dueling:
        import copy
        val = [copy.deepcopy(l) for l in adv[:-1]] + [dense(dense_specs[-1][0], 1, ACT_IDENTITY)]
    return RecurrentQ(lstm, val, adv)


class EpisodeReplayBuffer:
    """src/episode_replay.jl:3-95.  Index source: the engine's counter-based stream (the reference's MersenneTwister(0) stream is not
    reproducible offline): batch_size distinct episodes = the sum-tree sampler over unit priorities (uniform without replacement), then
    ep_start = 1 + floor(u * len), u = Philox(seed; slot, attempt 0x40000000, call)."""

    def __init__(self, obs_shape, max_size, batch_size, trace_length, max_len=100):
        self.max_size, self.batch_size, self.trace_length, self.max_len = int(max_size), int(batch_size), int(trace_length), int(max_len)
        self.obs_shape = tuple(obs_shape)
        self._curr_size, self._idx = 0, 0
        self._experience = [None] * self.max_size
        self.tree = SumTree(self.max_size)

    def add_episode(self, s, a, r, sp, done):
        """one whole episode: arrays of len steps (a 1-based)"""
        assert 1 <= len(a) <= self.max_len
        self._experience[self._idx] = (np.asarray(s, np.float32), np.asarray(a, np.int32), np.asarray(r, np.float32), np.asarray(sp, np.float32),
                                       np.asarray(done, np.uint8))
        self.tree.set_leaves([self._idx], [np.float32(1.0)])
        self._idx = (self._idx + 1) % self.max_size
        self._curr_size = min(self._curr_size + 1, self.max_size)

    def sample_indices(self, seed, call):
        assert self._curr_size >= self.batch_size
        idx, _ = self.tree.sample(self.batch_size, seed, call)
        u = sample_uniforms(seed, call, np.arange(self.batch_size, dtype=np.uint32), np.full(self.batch_size, 0x40000000, np.uint32))
        lens = np.array([len(self._experience[i][1]) for i in idx], np.int64)
        start = 1 + np.minimum((u * lens.astype(np.float32)).astype(np.float32).astype(np.int64), lens - 1)
        return idx, start

    def get_batch(self, idx, start):
        """-> s[T,B,...], a[T,B] (1-based; 1 where unfilled), r[T,B], sp[T,B,...], done[T,B] float, mask[T,B] int32   (:71-95)"""
        T, B = self.trace_length, self.batch_size
        s = np.zeros((T, B) + self.obs_shape, np.float32); sp = np.zeros_like(s)
        a = np.ones((T, B), np.int32); r = np.zeros((T, B), np.float32); d = np.zeros((T, B), np.float32); m = np.zeros((T, B), np.int32)
        for i, (e, st) in enumerate(zip(idx, start)):
            es, ea, er, esp, ed = self._experience[e]
            n = min(len(ea), T) - int(st) + 1              # for j = ep_start:min(length(ep), trace_length), copying ep[1], ep[2], ...
            for t in range(max(n, 0)):
                s[t, i], a[t, i], r[t, i], sp[t, i], d[t, i], m[t, i] = es[t], ea[t], er[t], esp[t], float(ed[t]), 1
        return s, a, r, sp, d, m


def forward_backward_recurrent(active_q, target_q, s, a, r, sp, done, mask, gamma, double_q=True, dtype=np.float32):
    """src/solver.jl:249-283.  a is 0-based here.  Returns dict(q[T,B,A], y[T,B], td[T,B], loss, grads, grad_norm, best_a)."""
    T, B = a.shape
    gam = dtype(np.float32(gamma))
    active_q.reset(); target_q.reset()
    ys, bests, qps, tqs = [], [], [], []
    for t in range(T):
        qp = active_q(sp[t], dtype)
        tq = target_q(sp[t], dtype)
        y, best = q_targets_of(qp, tq, r[t], done[t], gam, double_q)
        ys.append(y); bests.append(best); qps.append(qp); tqs.append(tq)
    active_q.reset()
    qs, caches, tds, dqs = [], [], [], []
    loss = dtype(0)
    for t in range(T):
        q, c = active_q.step(sp[t] * 0 + s[t], dtype)
        q_sa = q[np.arange(B), a[t]]
        td = q_sa - ys[t]
        x = mask[t].astype(dtype) * td
        loss = loss + huber_loss(x).sum(dtype=dtype) / dtype(B)
        g = mask[t].astype(dtype) * np.clip(x, dtype(-1), dtype(1)) / dtype(B) / dtype(T)
        dq = np.zeros_like(q); dq[np.arange(B), a[t]] = g
        qs.append(q); caches.append(c); tds.append(td); dqs.append(dq)
    loss = loss / dtype(T)
    grads = active_q.backward_seq(dqs, caches)
    trainable = [gr for k, gr in enumerate(grads) if k not in (3, 4)]         # state0 gets no gradient in the reference (module docstring)
    return dict(q=np.stack(qs), q_online_sp=np.stack(qps), q_target_sp=np.stack(tqs), y=np.stack(ys), td=np.stack(tds), best_a=np.stack(bests),
                loss=loss, grads=grads, grad_norm=globalnorm([gr.astype(np.float32) for gr in trainable]))


def batch_train_recurrent(active_q, target_q, optimizer, batch, gamma, double_q=True):
    s, a1, r, sp, done, mask = batch
    out = forward_backward_recurrent(active_q, target_q, s, a1 - 1, r, sp, done, mask, gamma, double_q, np.float32)
    grads = [None if k in (3, 4) else g for k, g in enumerate(out["grads"])]     # Adam skips `nothing` gradients
    optimizer.apply(active_q.params(), grads)
    return out["loss"], out["grad_norm"], out
