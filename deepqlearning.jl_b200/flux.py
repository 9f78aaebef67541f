"""Host-side descriptors with the reference's names: Chain / Dense / Conv / flattenbatch / DuelingNetwork.

These objects only *describe* a network and hold its parameters in Flux memory layout between
`solve` calls (Flux.params order; Dense weight (out,in) column-major == numpy (in,out); Conv weight
(kw,kh,cin,cout) column-major == numpy (cout,cin,kh,kw), true convolution).  They never compute a
forward pass - Q-values come from the engine (`dqn_q_values`, `dqn_train_step`).

Reference: src/dueling.jl:2-58 (DuelingNetwork, create_dueling_network), src/helpers.jl:6-8,25-32."""
import copy

import numpy as np

from . import _capi

identity, relu, tanh, sigmoid = _capi.ACT_IDENTITY, _capi.ACT_RELU, _capi.ACT_TANH, _capi.ACT_SIGMOID


class Dense:
    def __init__(self, nin, nout, act=identity, rng=None):
        self.nin, self.nout, self.act = int(nin), int(nout), int(act)
        rng = rng or np.random.default_rng()
        lim = np.sqrt(6.0 / (nin + nout))                                   # Flux glorot_uniform
        self.weight = rng.uniform(-lim, lim, (nin, nout)).astype(np.float32)
        self.bias = np.zeros(nout, np.float32)

    def params(self):
        return [self.weight, self.bias]

    def desc(self):
        return dict(kind=_capi.LAYER_DENSE, act=self.act, in_=self.nin, out=self.nout)


class Conv:
    """Conv((kw, kh), cin => cout, act; stride) in Flux's argument order (first kernel extent is W)."""

    def __init__(self, k, cin, cout, act=identity, stride=1, rng=None):
        self.kw, self.kh = int(k[0]), int(k[1])
        self.cin, self.cout, self.act, self.stride = int(cin), int(cout), int(act), int(stride)
        rng = rng or np.random.default_rng()
        lim = np.sqrt(6.0 / ((cin + cout) * self.kh * self.kw))
        self.weight = rng.uniform(-lim, lim, (cout, cin, self.kh, self.kw)).astype(np.float32)
        self.bias = np.zeros(cout, np.float32)

    def params(self):
        return [self.weight, self.bias]

    def desc(self):
        return dict(kind=_capi.LAYER_CONV, act=self.act, in_=self.cin, out=self.cout, kh=self.kh, kw=self.kw, stride=self.stride)


class LSTM:
    """Flux.LSTM(in, out) = Recur(LSTMCell): Wi (4out, in), Wh (4out, out), b (4out) with the forget-gate slice at 1, state0 = (h0, c0)
    zeros (out, 1); numpy images (in, 4out), (out, 4out), (4out,), (1, out), (1, out); Flux.params order Wi, Wh, b, h0, c0."""

    def __init__(self, nin, nout, rng=None):
        self.nin, self.nout = int(nin), int(nout)
        rng = rng or np.random.default_rng()
        lim = np.sqrt(6.0 / (nin + 4 * nout))
        self.Wi = rng.uniform(-lim, lim, (nin, 4 * nout)).astype(np.float32)
        lim = np.sqrt(6.0 / (nout + 4 * nout))
        self.Wh = rng.uniform(-lim, lim, (nout, 4 * nout)).astype(np.float32)
        self.b = np.zeros(4 * nout, np.float32)
        self.b[nout:2 * nout] = 1.0
        self.h0 = np.zeros((1, nout), np.float32)
        self.c0 = np.zeros((1, nout), np.float32)

    def params(self):
        return [self.Wi, self.Wh, self.b, self.h0, self.c0]

    def desc(self):
        return dict(kind=_capi.LAYER_LSTM, act=identity, in_=self.nin, out=self.nout)


class flattenbatch:
    """src/helpers.jl:6-8"""

    def params(self):
        return []

    def desc(self):
        return dict(kind=_capi.LAYER_FLATTEN, act=identity, in_=0, out=0)


class Chain:
    def __init__(self, *layers):
        self.layers = list(layers)

    def params(self):
        return [p for l in self.layers for p in l.params()]

    def __iter__(self):
        return iter(self.layers)

    def __len__(self):
        return len(self.layers)


class DuelingNetwork:
    """src/dueling.jl:2-6; Flux.params order is base, val, adv (:13)."""

    def __init__(self, base, val, adv):
        self.base, self.val, self.adv = base, val, adv

    def params(self):
        return self.base.params() + self.val.params() + self.adv.params()

    def __iter__(self):        # src/dueling.jl:19-30
        return iter(self.base.layers + self.val.layers + self.adv.layers)


def create_dueling_network(m, rng=None):
    """src/dueling.jl:36-58."""
    n = len(m.layers)
    duel_layer = -1
    for i in range(1, n + 1):
        if not isinstance(m.layers[n - i], Dense):
            duel_layer = n - i + 1
            break
        elif i == n:
            duel_layer = 0
    err = "DeepQLearningError: the qnetwork provided is incompatible with dueling"
    if duel_layer == -1 or duel_layer == n:
        raise ValueError(err)
    tail = m.layers[duel_layer:]
    val = Chain(*[copy.deepcopy(l) for l in tail[:-1]], Dense(tail[-1].nin, 1, rng=rng))
    adv = Chain(*[copy.deepcopy(l) for l in tail])
    base = Chain(*[copy.deepcopy(l) for l in m.layers[:duel_layer]])
    return DuelingNetwork(base, val, adv)


def isrecurrent(m):
    """src/helpers.jl:25-32"""
    return any(isinstance(l, LSTM) for l in m)


def flat_params(net):
    ps = net.params()
    return np.concatenate([p.ravel() for p in ps]).astype(np.float32) if ps else np.zeros(0, np.float32)


def load_flat_params(net, flat):
    """Flux.loadparams!"""
    o = 0
    for p in net.params():
        p[...] = np.asarray(flat[o:o + p.size], np.float32).reshape(p.shape)
        o += p.size
    assert o == len(flat)
