"""Mint the golden vectors under tests/golden/ from the oracle (run from the repo root:
`python tests/golden/make_golden.py`).  The reference holds no fixture for this path and cannot be executed
here (no Julia), so these pin the *restatement*: inputs + the oracle's fp32 and fp64 outputs of one
batch_train! step, for three small configurations.  Parity stays "unpinned" against the reference itself."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as O                      # noqa: E402
import util                             # noqa: E402

SEED = 2                                # sampler key
CASES = [("c1_gridworld", True, True), ("testmdp", True, True), ("conv_small", True, True), ("conv_small", False, False)]


def make(name, dueling, double_q):
    spec = util.SPECS[name]
    net = util.make_oracle_net(spec, dueling, seed=11)
    tgt = util.perturbed_copy(net, seed=12)
    buf = util.make_oracle_replay(spec)
    n = spec["N"] - 7
    s, a, r, sp, done = util.random_transitions(spec, n, seed=13)
    buf.add_batch(s, a, r, sp, done, np.abs(r))
    idx, attempts = buf.tree.sample(spec["B"], SEED, 0)
    theta0, theta_t = O.flat_params(net), O.flat_params(tgt)
    tree_total0 = np.float32(buf.tree.total)
    sb, ab, rb, spb, db, _, w = buf.get_batch(idx, total="tree", dequant=util.dequant)
    out64 = O.forward_backward(net, tgt, sb, ab - 1, rb, spb, db, w, 0.99, double_q, np.float64)
    opt = O.Adam(spec["lr"])
    loss, gn, out = O.batch_train(net, tgt, opt, buf, idx, 0.99, double_q, True, total="tree", dequant=util.dequant)
    g = dict(theta0=theta0, theta_t=theta_t, s=s, a=a, r=r, sp=sp, done=done, idx=idx, w=out["w"],
             q=out["q"], q_online_sp=out["q_online_sp"], q_target_sp=out["q_target_sp"], best_a=out["best_a"], y=out["y"], td=out["td"],
             loss=np.float32(loss), grad_norm=np.float32(gn), grads=np.concatenate([x.ravel() for x in out["grads"]]).astype(np.float32),
             theta1=O.flat_params(net), prio1=buf._priorities.copy(), tree_total=tree_total0,
             q64=out64["q"], td64=out64["td"], loss64=np.float64(out64["loss"]),
             grads64=np.concatenate([x.ravel() for x in out64["grads"]]))
    fn = os.path.join(ROOT, "tests", "golden", f"{name}_d{int(dueling)}q{int(double_q)}.npz")
    np.savez_compressed(fn, **g)
    print(fn, os.path.getsize(fn))


if __name__ == "__main__":
    for c in CASES:
        make(*c)
