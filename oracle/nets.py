"""Q-network restatement: Chain / Dense / Conv / flattenbatch / DuelingNetwork (TEST INFRASTRUCTURE).

Array convention: a Julia array of size (d1,...,dk) (column-major) is held as a C-ordered numpy
array of shape (dk,...,d1) - the memory image is identical.  Hence
    Flux Dense.weight (out,in)          <-> numpy (in,out)
    Flux Conv.weight  (kw,kh,cin,cout)  <-> numpy (cout,cin,kh,kw)
    data (W,H,C,N)                      <-> numpy (N,C,H,W)
    activations (feat,B)                <-> numpy (B,feat)
and `flat_params` is exactly the concatenation of the Flux.params arrays as they lie in memory.

Reference: src/dueling.jl:2-13 (forward), :36-58 (create_dueling_network), src/helpers.jl:6-8
(flattenbatch).  Flux semantics (Dense = sigma.(W*x .+ b); Conv = true convolution with flipped
kernel, no padding) are restated from SURVEY.md App. B.1/B.2 (Flux 0.14 / NNlib, source not
available offline).
"""
import copy
import numpy as np

ACT_IDENTITY, ACT_RELU, ACT_TANH, ACT_SIGMOID = 0, 1, 2, 3


def _act(z, act):
    if act == ACT_IDENTITY:
        return z
    if act == ACT_RELU:
        return np.maximum(z, z.dtype.type(0))
    if act == ACT_TANH:
        return np.tanh(z)
    if act == ACT_SIGMOID:
        one = z.dtype.type(1)
        return one / (one + np.exp(-z))
    raise ValueError(act)


def _dact(y, act):
    """sigma'(z) expressed through the output y = sigma(z)."""
    one = y.dtype.type(1)
    if act == ACT_IDENTITY:
        return np.ones_like(y)
    if act == ACT_RELU:
        return (y > 0).astype(y.dtype)
    if act == ACT_TANH:
        return one - y * y
    if act == ACT_SIGMOID:
        return y * (one - y)
    raise ValueError(act)


def _dact_of(layer, y):
    """sigma'(z) of a layer.  `layer.mask_override` (a 0/1 array shaped like y, set by a parity test) replaces the ReLU mask y > 0:
    a unit whose pre-activation is within rounding distance of zero flips its sub-gradient between any two fp32 evaluations, so the
    gradient comparison takes the mask from the implementation under test and checks everything else to full tolerance."""
    m = getattr(layer, "mask_override", None)
    if m is not None and layer.act == ACT_RELU:
        return np.asarray(m).reshape(y.shape).astype(y.dtype)
    return _dact(y, layer.act)


class Dense:
    def __init__(self, nin, nout, act=ACT_IDENTITY, weight=None, bias=None):
        self.nin, self.nout, self.act = int(nin), int(nout), int(act)
        self.weight = np.zeros((nin, nout), np.float32) if weight is None else np.asarray(weight, np.float32)
        self.bias = np.zeros((nout,), np.float32) if bias is None else np.asarray(bias, np.float32)
        assert self.weight.shape == (nin, nout) and self.bias.shape == (nout,)

    def params(self):
        return [self.weight, self.bias]

    def forward(self, x, dtype):
        y = _act(x @ self.weight.astype(dtype) + self.bias.astype(dtype), self.act)
        return y, (x, y)

    def backward(self, dy, cache):
        x, y = cache
        delta = dy * _dact_of(self, y)
        dw = x.T @ delta
        db = delta.sum(axis=0)
        dx = delta @ self.weight.astype(dy.dtype).T
        return dx, [dw, db]


def _im2col(x, kh, kw, s):
    win = np.lib.stride_tricks.sliding_window_view(x, (kh, kw), axis=(2, 3))[:, :, ::s, ::s]
    n, c, oh, ow = win.shape[:4]
    return np.ascontiguousarray(win.transpose(0, 2, 3, 1, 4, 5)).reshape(n * oh * ow, c * kh * kw), (n, oh, ow)


class Conv:
    """Flux Conv((kw,kh), cin=>cout, act; stride, pad=0): true convolution (kernel flipped)."""

    def __init__(self, kh, kw, cin, cout, stride=1, act=ACT_IDENTITY, weight=None, bias=None):
        self.kh, self.kw, self.cin, self.cout, self.stride, self.act = int(kh), int(kw), int(cin), int(cout), int(stride), int(act)
        self.weight = np.zeros((cout, cin, kh, kw), np.float32) if weight is None else np.asarray(weight, np.float32)
        self.bias = np.zeros((cout,), np.float32) if bias is None else np.asarray(bias, np.float32)
        assert self.weight.shape == (cout, cin, kh, kw)

    def params(self):
        return [self.weight, self.bias]

    def out_hw(self, h, w):
        return (h - self.kh) // self.stride + 1, (w - self.kw) // self.stride + 1

    def forward(self, x, dtype):
        cols, (n, oh, ow) = _im2col(x, self.kh, self.kw, self.stride)
        wf = self.weight[:, :, ::-1, ::-1].astype(dtype).reshape(self.cout, -1)   # flipped => cross-correlation
        z = cols @ wf.T + self.bias.astype(dtype)
        y = _act(z, self.act).reshape(n, oh, ow, self.cout).transpose(0, 3, 1, 2)
        return np.ascontiguousarray(y), (x.shape, cols, y)

    def backward(self, dy, cache):
        xshape, cols, y = cache
        n, c, h, w = xshape
        oh, ow = y.shape[2], y.shape[3]
        delta = (dy * _dact_of(self, y)).transpose(0, 2, 3, 1).reshape(-1, self.cout)
        dwf = (cols.T @ delta).T.reshape(self.cout, self.cin, self.kh, self.kw)
        dw = np.ascontiguousarray(dwf[:, :, ::-1, ::-1])
        db = delta.sum(axis=0)
        wf = self.weight[:, :, ::-1, ::-1].astype(dy.dtype).reshape(self.cout, -1)
        dcols = (delta @ wf).reshape(n, oh, ow, c, self.kh, self.kw)
        dx = np.zeros(xshape, dy.dtype)
        s = self.stride
        for j in range(self.kh):
            for i in range(self.kw):
                dx[:, :, j:j + s * oh:s, i:i + s * ow:s] += dcols[:, :, :, :, j, i].transpose(0, 3, 1, 2)
        return dx, [dw, db]


class Flatten:
    """flattenbatch (src/helpers.jl:6-8): reshape(x, (:, B)) - a memory identity."""

    def params(self):
        return []

    def forward(self, x, dtype):
        return x.reshape(x.shape[0], -1), x.shape

    def backward(self, dy, cache):
        return dy.reshape(cache), []


class Chain:
    def __init__(self, *layers):
        self.layers = list(layers)

    def params(self):
        return [p for l in self.layers for p in l.params()]

    def forward(self, x, dtype=np.float32):
        caches = []
        for l in self.layers:
            x, c = l.forward(x, dtype)
            caches.append(c)
        return x, caches

    def backward(self, dy, caches):
        grads = []
        for l, c in zip(reversed(self.layers), reversed(caches)):
            dy, g = l.backward(dy, c)
            grads = g + grads
        return dy, grads

    def __call__(self, x, dtype=np.float32):
        return self.forward(np.asarray(x, dtype), dtype)[0]


def _seq_sum_actions(a):
    """sum over the action axis in index order, one rounded add per action (Julia column reduce)."""
    m = a[:, 0].copy()
    for k in range(1, a.shape[1]):
        m = m + a[:, k]
    return m


class DuelingNetwork:
    """src/dueling.jl:2-11:  Q = val(x) .+ adv(x) .- mean(adv(x), dims=1),  x = base(inpt)."""

    def __init__(self, base, val, adv):
        self.base, self.val, self.adv = base, val, adv

    def params(self):   # Flux.@functor field order (src/dueling.jl:2-6,13)
        return self.base.params() + self.val.params() + self.adv.params()

    def forward(self, x, dtype=np.float32):
        xb, cb = self.base.forward(x, dtype)
        v, cv = self.val.forward(xb, dtype)
        a, ca = self.adv.forward(xb, dtype)
        mean = _seq_sum_actions(a) / dtype(a.shape[1])
        q = (v + a) - mean[:, None]
        return q, (cb, cv, ca, a.shape[1])

    def backward(self, dq, caches):
        cb, cv, ca, na = caches
        dv = _seq_sum_actions(dq)[:, None]
        da = dq - (dv / dq.dtype.type(na))
        dxv, gv = self.val.backward(dv, cv)
        dxa, ga = self.adv.backward(da, ca)
        dx, gb = self.base.backward(dxv + dxa, cb)
        return dx, gb + gv + ga

    def __call__(self, x, dtype=np.float32):
        return self.forward(np.asarray(x, dtype), dtype)[0]


def create_dueling_network(m):
    """src/dueling.jl:36-58.  The trailing run of Dense layers becomes the adv tower; the val tower is
    a copy of all of them but the last plus a fresh Dense(k, 1), k = input width of the last Dense."""
    n = len(m.layers)
    duel_layer = -1
    for i in range(1, n + 1):
        l = m.layers[n - i]
        if not isinstance(l, Dense):
            duel_layer = n - i + 1          # 1-based index of the last non-Dense layer
            break
        elif i == n:
            duel_layer = 0
    if duel_layer == -1 or duel_layer == n:
        raise ValueError("DeepQLearningError: the qnetwork provided is incompatible with dueling")
    tail = m.layers[duel_layer:]
    last = tail[-1]
    val = Chain(*[copy.deepcopy(l) for l in tail[:-1]], Dense(last.nin, 1))
    adv = Chain(*[copy.deepcopy(l) for l in tail])
    base = Chain(*[copy.deepcopy(l) for l in m.layers[:duel_layer]])
    return DuelingNetwork(base, val, adv)


def glorot_uniform_chain(net, rng):
    """Flux default init: W ~ U(-1,1)*sqrt(6/(fan_in+fan_out)), b = 0.  Uses numpy's
    generator, not Julia's - initial weights are inputs to the path, not part of it."""
    for p_owner in _layers_of(net):
        if isinstance(p_owner, Dense):
            lim = np.sqrt(6.0 / (p_owner.nin + p_owner.nout))
            p_owner.weight[...] = rng.uniform(-lim, lim, p_owner.weight.shape).astype(np.float32)
        elif isinstance(p_owner, Conv):
            fan_in = p_owner.cin * p_owner.kh * p_owner.kw
            fan_out = p_owner.cout * p_owner.kh * p_owner.kw
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            p_owner.weight[...] = rng.uniform(-lim, lim, p_owner.weight.shape).astype(np.float32)
    return net


def _layers_of(net):
    if isinstance(net, DuelingNetwork):
        return net.base.layers + net.val.layers + net.adv.layers
    return net.layers


def params_of(net):
    return net.params()


def num_params(net):
    return int(sum(p.size for p in net.params()))


def flat_params(net):
    ps = net.params()
    return np.concatenate([p.ravel() for p in ps]) if ps else np.zeros(0, np.float32)


def set_params(net, flat):
    flat = np.asarray(flat, np.float32)
    o = 0
    for p in net.params():
        p[...] = flat[o:o + p.size].reshape(p.shape)
        o += p.size
    assert o == flat.size
