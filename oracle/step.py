"""batch_train! restatement (TEST INFRASTRUCTURE) - src/solver.jl:191-236, numbered as SURVEY App. A.

    4-5  forward s' online/target, Double-Q argmax (first max)          solver.jl:209-216
    6    y = r + ((1-d)*gamma)*q'   three separately rounded fp32 ops    solver.jl:217
    7    td = Q(s)[a] - y                                                solver.jl:220-222
    8    L = sum(huber(w*td))/B                                          solver.jl:223-224, helpers.jl:14-19
    9    reverse pass (hand-derived here; torch autograd re-derives it in tests/)
    10   grad_norm = max |g|                                             helpers.jl:38-46
    11   Flux.Optimise.Adam, Float64 scalars, fp32 state                 solver.jl:66,228
    12   priorities <- (|td|+eps)^alpha                                   PER.jl:76-80

`dtype=np.float64` evaluates the same step in double precision (the distance of the fp32
restatement to it is reported next to every GPU comparison).
"""
import numpy as np


def huber_loss(x):
    """src/helpers.jl:14-19 (delta = 1)."""
    one = x.dtype.type(1)
    abserror = np.abs(x)
    quadratic = np.minimum(abserror, one)
    linear = abserror - quadratic
    return x.dtype.type(0.5) * quadratic * quadratic + linear


def globalnorm(grads):
    """src/helpers.jl:38-46: max over parameter arrays of max|g| (an infinity norm, not a 2-norm)."""
    g = np.float32(0)
    for a in grads:
        if a is None or a.size == 0:
            continue
        c = np.float32(np.max(np.abs(a)))
        g = c if c > g else g
    return g


def q_targets_of(q_online_sp, q_target_sp, r, done, gamma, double_q):
    """solver.jl:209-217.  Returns (y, best_a) with best_a 0-based."""
    dt = q_target_sp.dtype.type
    if double_q:
        best_a = np.argmax(q_online_sp, axis=1)              # first maximal index, as Julia argmax
        q_sp_max = q_target_sp[np.arange(q_target_sp.shape[0]), best_a]
    else:
        best_a = np.argmax(q_target_sp, axis=1)
        q_sp_max = q_target_sp.max(axis=1)
    y = r.astype(q_target_sp.dtype) + ((dt(1) - done.astype(q_target_sp.dtype)) * dt(gamma)) * q_sp_max
    return y, best_a


def forward_backward(active_q, target_q, s, a, r, sp, done, w, gamma, double_q=True, dtype=np.float32):
    """Steps 4-10.  `a` is 0-based here (the reference's CartesianIndex(a_i, i) with a_i 1-based)."""
    B = s.shape[0]
    s = np.asarray(s, dtype)
    sp = np.asarray(sp, dtype)
    gam = dtype(np.float32(gamma))
    qp = active_q(sp, dtype)
    tq = target_q(sp, dtype)
    y, best_a = q_targets_of(qp, tq, np.asarray(r, np.float32), np.asarray(done, np.float32), gam, double_q)
    q, caches = active_q.forward(s, dtype)
    q_sa = q[np.arange(B), a]
    td = q_sa - y
    x = w.astype(dtype) * td
    loss = huber_loss(x).sum(dtype=dtype) / dtype(B)
    g = w.astype(dtype) * np.clip(x, dtype(-1), dtype(1)) / dtype(B)
    dq = np.zeros_like(q)
    dq[np.arange(B), a] = g
    _, grads = active_q.backward(dq, caches)
    return dict(q=q, q_online_sp=qp, q_target_sp=tq, best_a=best_a, y=y, td=td, loss=loss, g=g,
                grads=grads, grad_norm=globalnorm([gr.astype(np.float32) for gr in grads]))


class Adam:
    """Flux.Optimise.Adam: eta/beta/epsilon and the running beta powers are Float64,
    the moments are Float32 arrays; every element is computed in Float64 and rounded on store."""

    def __init__(self, eta, beta=(0.9, 0.999), epsilon=1e-8):
        self.eta = float(np.float32(eta))       # learning_rate::Float32 widened (solver.jl:3,66)
        self.beta = (float(beta[0]), float(beta[1]))
        self.epsilon = float(epsilon)
        self.state = {}

    def apply(self, params, grads):
        for k, (x, d) in enumerate(zip(params, grads)):
            if d is None:
                continue
            if k not in self.state:
                self.state[k] = [np.zeros_like(x, np.float32), np.zeros_like(x, np.float32), [self.beta[0], self.beta[1]]]
            mt, vt, bp = self.state[k]
            b1, b2 = self.beta
            d64 = d.astype(np.float32).astype(np.float64)
            mt[...] = (b1 * mt.astype(np.float64) + (1.0 - b1) * d64).astype(np.float32)
            vt[...] = (b2 * vt.astype(np.float64) + (1.0 - b2) * d64 * d64).astype(np.float32)
            upd = (mt.astype(np.float64) / (1.0 - bp[0]) / (np.sqrt(vt.astype(np.float64) / (1.0 - bp[1])) + self.epsilon) * self.eta).astype(np.float32)
            bp[0] *= b1
            bp[1] *= b2
            x[...] = x - upd


def batch_train(active_q, target_q, optimizer, replay, idx, gamma, double_q=True, prioritized_replay=True,
                total="pairwise", dequant=None):
    """One full step on the given sampled indices (0-based).  Returns (loss, grad_norm, info)."""
    s, a1, r, sp, done, idx, w = replay.get_batch(idx, total=total, dequant=dequant)
    out = forward_backward(active_q, target_q, s, a1 - 1, r, sp, done, w, gamma, double_q, np.float32)
    optimizer.apply(active_q.params(), out["grads"])
    if prioritized_replay:
        replay.update_priorities(idx, out["td"])
    out["w"] = w
    return out["loss"], out["grad_norm"], out
