"""Shared helpers: one network/replay spec -> (oracle objects, engine config) so that the GPU parity tests,
the golden-vector generator and smoke() all drive both sides from identical inputs."""
import copy

import numpy as np

import oracle as O
from oracle.synthetic import synthetic_transitions

# name -> dict(layers=[...], obs=(C,H,W)|(d,), nA, B, N, u8)
#   layer tuples: ("dense", in, out, act) | ("conv", k, cin, cout, stride, act) | ("flatten",)
SPECS = {
    # README example (README.md:38): Dense(2,32) -> Dense(32,4), identity activations, dueling => two towers
    "c1_gridworld": dict(layers=[("dense", 2, 32, 0), ("dense", 32, 4, 0)], obs=(2,), nA=4, B=32, N=1000, u8=False, lr=5e-3),
    # reference test-suite network (test/runtests.jl:98) on TestMDP((5,5),4,6): flatten -> Dense(100,8,tanh) -> Dense(8,4)
    "testmdp": dict(layers=[("flatten",), ("dense", 100, 8, 2), ("dense", 8, 4, 0)], obs=(4, 5, 5), nA=4, B=32, N=500, u8=False, lr=5e-3),
    # small conv trunk, u8 store, odd sizes
    "conv_small": dict(layers=[("conv", 4, 4, 8, 2, 1), ("conv", 3, 8, 12, 1, 1), ("flatten",), ("dense", 12 * 3 * 3, 20, 1), ("dense", 20, 5, 0)],
                       obs=(4, 12, 12), nA=5, B=24, N=300, u8=True, lr=1e-3),
    # smooth (tanh) networks, every contraction large enough for the tensor-core path: gradient parity without ReLU sub-gradient flips
    "mlp_tanh": dict(layers=[("dense", 128, 256, 2), ("dense", 256, 256, 2), ("dense", 256, 16, 0)], obs=(128,), nA=16, B=256, N=2048, u8=False, lr=1e-4),
    "conv_tanh": dict(layers=[("conv", 4, 4, 32, 2, 2), ("conv", 3, 32, 32, 1, 2), ("flatten",), ("dense", 32 * 7 * 7, 64, 2), ("dense", 64, 5, 0)],
                      obs=(4, 20, 20), nA=5, B=64, N=512, u8=True, lr=1e-3),
    # config 2 (BASELINE.json): 128 -> 3x256 -> 16
    "c2_mlp": dict(layers=[("dense", 128, 256, 1), ("dense", 256, 256, 1), ("dense", 256, 256, 1), ("dense", 256, 16, 0)],
                   obs=(128,), nA=16, B=256, N=4096, u8=False, lr=1e-4),
    # config 3 (BASELINE.json): Nature-DQN trunk + dueling, Atari-shaped u8 observations
    "c3_conv": dict(layers=[("conv", 8, 4, 32, 4, 1), ("conv", 4, 32, 64, 2, 1), ("conv", 3, 64, 64, 1, 1), ("flatten",),
                            ("dense", 3136, 512, 1), ("dense", 512, 6, 0)], obs=(4, 84, 84), nA=6, B=256, N=2048, u8=True, lr=1e-4),
}


def oracle_chain(spec):
    ls = []
    for l in spec["layers"]:
        if l[0] == "dense":
            ls.append(O.Dense(l[1], l[2], l[3]))
        elif l[0] == "conv":
            ls.append(O.Conv(l[1], l[1], l[2], l[3], l[4], l[5]))
        else:
            ls.append(O.Flatten())
    return O.Chain(*ls)


def layer_descs(spec):
    """dicts for dqn_layer_t (kind: 0 dense, 1 conv, 2 flatten)."""
    out = []
    for l in spec["layers"]:
        if l[0] == "dense":
            out.append(dict(kind=0, act=l[3], in_=l[1], out=l[2]))
        elif l[0] == "conv":
            out.append(dict(kind=1, act=l[5], in_=l[2], out=l[3], kh=l[1], kw=l[1], stride=l[4]))
        else:
            out.append(dict(kind=2, act=0, in_=0, out=0))
    return out


def make_oracle_net(spec, dueling, seed):
    rng = np.random.default_rng(seed)
    net = oracle_chain(spec)
    if dueling:
        net = O.create_dueling_network(net)
    O.glorot_uniform_chain(net, rng)
    for p in net.params():          # biases away from zero so that their gradients / layout are exercised
        if p.ndim == 1:
            p[...] = rng.normal(0, 0.05, p.shape).astype(np.float32)
    return net


def perturbed_copy(net, seed, scale=0.02):
    rng = np.random.default_rng(seed)
    t = copy.deepcopy(net)
    for p in t.params():
        p += rng.normal(0, scale, p.shape).astype(np.float32)
    return t


def random_transitions(spec, n, seed):
    rng = np.random.default_rng(seed)
    shape = tuple(spec["obs"])
    if spec["u8"]:
        s = rng.integers(0, 256, (n,) + shape, dtype=np.uint8)
        sp = rng.integers(0, 256, (n,) + shape, dtype=np.uint8)
    else:
        s = rng.normal(0, 1, (n,) + shape).astype(np.float32)
        sp = rng.normal(0, 1, (n,) + shape).astype(np.float32)
    a = rng.integers(1, spec["nA"] + 1, n).astype(np.int32)
    r = rng.uniform(-1, 1, n).astype(np.float32)
    done = (rng.uniform(size=n) < 0.1).astype(np.uint8)
    return s, a, r, sp, done


def dequant(x):
    return x.astype(np.float32) / np.float32(255) if x.dtype == np.uint8 else x


def make_oracle_replay(spec, alpha=0.6, beta=0.4, eps=1e-3):
    return O.PrioritizedReplayBuffer(tuple(spec["obs"]), spec["N"], spec["B"], alpha, beta, eps, obs_dtype=np.uint8 if spec["u8"] else np.float32)


def relerr(a, b):
    """normwise relative error max|a-b| / max|b| (the tolerance the north_star states is 1e-5 on Q-values)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / den
