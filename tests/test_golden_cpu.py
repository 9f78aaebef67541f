"""The oracle reproduces the committed golden vectors (regression pin of the restatement; no GPU)."""
import glob
import os

import numpy as np
import pytest

import oracle as O
import util

GOLD = sorted(p for p in glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")) if not os.path.basename(p).startswith("ref_"))


def rebuild(g, name, dueling):
    spec = util.SPECS[name]
    net = util.make_oracle_net(spec, dueling, seed=0)
    tgt = util.make_oracle_net(spec, dueling, seed=0)
    O.set_params(net, g["theta0"])
    O.set_params(tgt, g["theta_t"])
    buf = util.make_oracle_replay(spec)
    buf.add_batch(g["s"], g["a"], g["r"], g["sp"], g["done"], np.abs(g["r"]))
    return spec, net, tgt, buf


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p) for p in GOLD])
def test_oracle_reproduces_golden(path):
    g = np.load(path)
    base = os.path.basename(path)[:-4]
    name, flags = base.rsplit("_", 1)
    dueling, double_q = flags[1] == "1", flags[3] == "1"
    spec, net, tgt, buf = rebuild(g, name, dueling)
    idx, _ = buf.tree.sample(spec["B"], 2, 0)
    assert np.array_equal(idx, g["idx"])                                   # integer path: bit-exact
    assert np.float32(buf.tree.total) == g["tree_total"]
    opt = O.Adam(spec["lr"])
    loss, gn, out = O.batch_train(net, tgt, opt, buf, idx, 0.99, double_q, True, total="tree", dequant=util.dequant)
    assert np.array_equal(out["best_a"], g["best_a"])
    for k in ("q", "td", "y", "w"):
        assert util.relerr(out[k], g[k]) < 1e-6, k
    assert abs(loss - g["loss"]) <= 1e-6 * abs(g["loss"])
    grads = np.concatenate([x.ravel() for x in out["grads"]])
    assert util.relerr(grads, g["grads"]) < 1e-5
    assert util.relerr(grads, g["grads64"]) < 1e-4                          # fp32 restatement vs its fp64 evaluation
    assert util.relerr(out["q"], g["q64"]) < 1e-5
    assert np.abs(O.flat_params(net) - g["theta1"]).max() < 1e-6
    np.testing.assert_allclose(buf._priorities, g["prio1"], rtol=1e-5)


REF = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "ref_*.npz")))


@pytest.mark.skipif(not REF, reason="no reference-minted fixtures (run oracle/julia/mint_fixtures.jl where Julia + DeepQLearning.jl are "
                                    "installed; until then the oracle is 'parity unpinned', DESIGN.md section 5)")
@pytest.mark.parametrize("path", REF, ids=[os.path.basename(p) for p in REF])
def test_oracle_against_reference_fixtures(path):
    """tests/golden/ref_<case>.npz are outputs of the REFERENCE's own batch_train! statements (Flux / Zygote / StatsBase) on the inputs of
    tests/golden/<case>.npz.  The restatement must reproduce them: indices-dependent integers exactly, floats to fp32 round-off."""
    ref = np.load(path)
    base = os.path.basename(path)[4:-4]
    g = np.load(os.path.join(os.path.dirname(path), base + ".npz"))
    name, flags = base.rsplit("_", 1)
    dueling, double_q = flags[1] == "1", flags[3] == "1"
    spec, net, tgt, buf = rebuild(g, name, dueling)
    opt = O.Adam(spec["lr"])
    loss, gn, out = O.batch_train(net, tgt, opt, buf, g["idx"], 0.99, double_q, True, total="pairwise", dequant=util.dequant)
    assert np.array_equal(out["best_a"], ref["best_a"])
    for k in ("q", "td", "y", "w"):
        assert util.relerr(out[k], ref[k]) < 1e-5, k
    assert abs(loss - float(ref["loss"][0])) <= 1e-5 * abs(float(ref["loss"][0]))
    grads = np.concatenate([x.ravel() for x in out["grads"]])
    assert util.relerr(grads, ref["grads"]) < 1e-4
    assert abs(gn - float(ref["grad_norm"][0])) <= 1e-4 * float(ref["grad_norm"][0])
    assert np.abs(O.flat_params(net) - ref["theta1"]).max() < 1e-6
    np.testing.assert_allclose(buf._priorities, ref["prio1"], rtol=1e-5)
