"""Flux-equivalent CPU restatement of batch_train! on torch-CPU (TEST / BASELINE INFRASTRUCTURE).

This is the CPU arm of bench.py (`cpu_baseline`, `--impl reference`): Julia/Flux cannot run in this image, so
the reference's CPU path is timed as a restatement that keeps the reference's algorithmic structure -
  * sampling as src/prioritized_experience_replay.jl:82-104: O(N) priority slice copy + Weights sum, A-ExpJ
    weighted sampling without replacement, per-sample gather loop, second O(N) sum for the IS weights;
  * three fp32 forwards (online s', target s', online s; the adv tower evaluated twice as src/dueling.jl:10),
    reverse-mode autograd, Flux-style Adam, update_priorities! -
on torch-CPU (MKL/oneDNN) with all host threads.  Label: "Flux-equivalent CPU restatement (torch-CPU)"."""
import os
import time

import numpy as np
import torch
import torch.nn.functional as F

from .replay import efraimidis_aexpj_wsample_norep


class TorchDuelingNet:
    """Parameters as torch tensors in Flux layout; forward follows src/dueling.jl:8-11."""

    def __init__(self, layers, dueling, rng):
        self.layers, self.dueling = layers, dueling
        self.trunk, self.val, self.adv = [], [], []
        dense = [l for l in layers if l[0] == "dense"]
        for l in layers:
            if l[0] == "conv":
                k, cin, cout = l[1], l[2], l[3]
                lim = np.sqrt(6.0 / ((cin + cout) * k * k))
                self.trunk.append(dict(w=torch.tensor(rng.uniform(-lim, lim, (cout, cin, k, k)).astype(np.float32), requires_grad=True),
                                       b=torch.zeros(cout, requires_grad=True), stride=l[4], act=l[5]))
        def mk(nin, nout, act):
            lim = np.sqrt(6.0 / (nin + nout))
            return dict(w=torch.tensor(rng.uniform(-lim, lim, (nin, nout)).astype(np.float32), requires_grad=True),
                        b=torch.zeros(nout, requires_grad=True), act=act)
        self.adv = [mk(l[1], l[2], l[3]) for l in dense]
        if dueling:
            self.val = [mk(l[1], l[2], l[3]) for l in dense[:-1]] + [mk(dense[-1][1], 1, 0)]

    def params(self):
        ps = []
        for l in self.trunk + self.val + self.adv:
            ps += [l["w"], l["b"]]
        return ps

    @staticmethod
    def _act(x, a):
        return [lambda z: z, torch.relu, torch.tanh, torch.sigmoid][a](x)

    def _tower(self, tower, x):
        for l in tower:
            x = self._act(x @ l["w"] + l["b"], l["act"])
        return x

    def __call__(self, x):
        for l in self.trunk:
            x = self._act(F.conv2d(x, torch.flip(l["w"], dims=(2, 3)), l["b"], stride=l["stride"]), l["act"])
        x = x.reshape(x.shape[0], -1)
        if not self.dueling:
            return self._tower(self.adv, x)
        return self._tower(self.val, x) + self._tower(self.adv, x) - self._tower(self.adv, x).mean(dim=1, keepdim=True)   # adv twice, dueling.jl:10

    def load_from(self, other):
        with torch.no_grad():
            for p, q in zip(self.params(), other.params()):
                p.copy_(q)


class CpuBaseline:
    def __init__(self, layers, obs_shape, n_actions, B, N, u8, store_rows=4096, lr=1e-4, gamma=0.99, seed=0, threads=None):
        self.threads = threads or len(os.sched_getaffinity(0))
        torch.set_num_threads(self.threads)
        rng = np.random.default_rng(seed)
        self.rng = rng
        self.B, self.N, self.u8, self.gamma, self.lr = B, N, u8, np.float32(gamma), float(np.float32(lr))
        self.net = TorchDuelingNet(layers, True, rng)
        self.tgt = TorchDuelingNet(layers, True, rng)
        self.tgt.load_from(self.net)
        self.store_rows = min(store_rows, N)                    # bounded observation store: transition i reads row i % store_rows
        shape = (self.store_rows,) + tuple(obs_shape)
        if u8:
            self.s = rng.integers(0, 256, shape, dtype=np.uint8)
            self.sp = rng.integers(0, 256, shape, dtype=np.uint8)
        else:
            self.s = rng.normal(0, 1, shape).astype(np.float32)
            self.sp = rng.normal(0, 1, shape).astype(np.float32)
        self.a = rng.integers(0, n_actions, N)
        self.r = rng.uniform(-1, 1, N).astype(np.float32)
        self.done = (rng.uniform(size=N) < 0.01).astype(np.float32)
        self.alpha, self.beta, self.eps = np.float32(0.6), np.float32(0.4), np.float32(1e-3)
        self.prio = (np.abs(self.r) + self.eps) ** self.alpha
        self.curr_size = N
        self.s_batch = np.zeros((B,) + tuple(obs_shape), np.float32)
        self.sp_batch = np.zeros((B,) + tuple(obs_shape), np.float32)
        self.m = [torch.zeros_like(p) for p in self.net.params()]
        self.v = [torch.zeros_like(p) for p in self.net.params()]
        self.bp = [0.9, 0.999]

    def step(self):
        B, n = self.B, self.curr_size
        w = self.prio[:n].copy()                                 # PER.jl:85 slice copy
        _ = w.sum()                                              # Weights(...)
        idx = efraimidis_aexpj_wsample_norep(self.rng, w, B)
        for i, k in enumerate(idx):                              # PER.jl:91-100 per-sample copy loop
            row = k % self.store_rows
            if self.u8:
                np.divide(self.s[row], np.float32(255), out=self.s_batch[i], dtype=np.float32)
                np.divide(self.sp[row], np.float32(255), out=self.sp_batch[i], dtype=np.float32)
            else:
                self.s_batch[i] = self.s[row]
                self.sp_batch[i] = self.sp[row]
        a_b, r_b, d_b = self.a[idx], self.r[idx], self.done[idx]
        p = self.prio[idx] / self.prio[:n].sum()                 # PER.jl:101 second O(N) sum
        isw = torch.from_numpy(((n * p) ** (-self.beta)).astype(np.float32))
        s, sp = torch.from_numpy(self.s_batch), torch.from_numpy(self.sp_batch)
        with torch.no_grad():                                    # solver.jl:209-217
            qp = self.net(sp)
            tq = self.tgt(sp)
            best = qp.argmax(dim=1)
            qmax = tq[torch.arange(B), best]
            y = torch.from_numpy(r_b) + (1 - torch.from_numpy(d_b)) * float(self.gamma) * qmax
        q = self.net(s)                                          # solver.jl:219-225
        td = q[torch.arange(B), torch.from_numpy(a_b)] - y
        x = isw * td
        ax = x.abs()
        quad = torch.clamp(ax, max=1.0)
        loss = (0.5 * quad * quad + (ax - quad)).sum() / B
        ps = self.net.params()
        grads = torch.autograd.grad(loss, ps)
        gnorm = max(float(g.abs().max()) for g in grads)         # helpers.jl:38-46
        with torch.no_grad():                                    # Flux.Optimise.Adam
            for p_, g, m, v in zip(ps, grads, self.m, self.v):
                m.mul_(0.9).add_(g, alpha=0.1)
                v.mul_(0.999).addcmul_(g, g, value=0.001)
                p_.sub_(m / (1 - self.bp[0]) / ((v / (1 - self.bp[1])).sqrt() + 1e-8) * self.lr)
            self.bp[0] *= 0.9
            self.bp[1] *= 0.999
        self.prio[idx] = (np.abs(td.detach().numpy()) + self.eps) ** self.alpha     # PER.jl:76-80
        return float(loss.detach()), gnorm

    def time_steps(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            self.step()
            ts.append(time.perf_counter() - t0)
        return ts
