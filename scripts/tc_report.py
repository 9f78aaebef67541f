"""Numerical report: fp32 CUDA-core path vs tcgen05 3xTF32 path vs the fp32 / fp64 oracle on config 3 (B=256)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dqn_b200 as lib       # noqa: E402
import oracle as O           # noqa: E402
import util                  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c3_conv"
spec = dict(util.SPECS[name])
net = util.make_oracle_net(spec, True, seed=21)
tgt = util.perturbed_copy(net, seed=22)
buf = util.make_oracle_replay(spec)
n = min(spec["N"], 600)
s, a, r, sp, done = util.random_transitions(spec, n, seed=23)
buf.add_batch(s, a, r, sp, done, np.abs(r))
idx, _ = buf.tree.sample(spec["B"], 2, 0)
sb, ab, rb, spb, db, _, w = buf.get_batch(idx, total="tree", dequant=util.dequant)
t0 = time.time()
o32 = O.forward_backward(net, tgt, sb, ab - 1, rb, spb, db, w, 0.99, True, np.float32)
o64 = O.forward_backward(net, tgt, sb, ab - 1, rb, spb, db, w, 0.99, True, np.float64)
print(f"oracle fp32+fp64 took {time.time() - t0:.1f}s")
g32 = np.concatenate([x.ravel() for x in o32["grads"]])
g64 = np.concatenate([x.ravel() for x in o64["grads"]])
print(f"oracle32 vs fp64: q {util.relerr(o32['q'], o64['q']):.2e} td {util.relerr(o32['td'], o64['td']):.2e} grads {util.relerr(g32, g64):.2e}")
for mode, label in ((0, "fp32-simt"), (1, "3xtf32-tcgen05")):
    cfg = lib.make_config(util.layer_descs(spec), tuple(reversed(spec["obs"])), spec["nA"], obs_dtype="u8" if spec["u8"] else "f32",
                          batch_size=spec["B"], buffer_size=spec["N"], learning_rate=spec["lr"], discount=0.99, seed=2, math_mode=mode)
    eng = lib.Engine(cfg)
    eng.set_params(O.flat_params(net), 0)
    eng.set_params(O.flat_params(tgt), 1)
    eng.replay_add(s, a, r, sp, done, np.abs(r))
    loss, gn = eng.train_step()
    assert np.array_equal(eng.last_indices(), idx)
    g = eng.grads()
    scale = np.abs(o64["q"]).max()
    print(f"[{label}] loss {loss:.7f} (oracle {o32['loss']:.7f}) gn {gn:.6e} (oracle {o32['grad_norm']:.6e})")
    print(f"[{label}] vs fp64: q {np.abs(eng.q(0) - o64['q']).max() / scale:.2e} qsp_on {np.abs(eng.q(1) - o64['q_online_sp']).max() / scale:.2e} "
          f"qsp_tg {np.abs(eng.q(2) - o64['q_target_sp']).max() / scale:.2e} td {np.abs(eng.td() - o64['td']).max() / scale:.2e} grads {util.relerr(g, g64):.2e}")
    o = 0
    for k, arr in enumerate(o64["grads"]):
        sl = slice(o, o + arr.size)
        print(f"    grad[{k}] shape {arr.shape}: rel {util.relerr(g[sl], g64[sl]):.2e}  (oracle32 {util.relerr(g32[sl], g64[sl]):.2e})")
        o += arr.size
    eng.set_profiling(1)
    for _ in range(5):
        eng.train_step_async()
    eng.sync()
    prof = eng.get_profile()
    eng.set_profiling(0)
    tot = sum(k["ms"] for k in prof)
    print(f"[{label}] eager per-kernel total {tot:.3f} ms:")
    for k in prof:
        tf = k["flops"] / (k["ms"] * 1e-3) / 1e12 if k["flops"] else 0
        print(f"    {k['name']:24s} {k['ms']*1e3:9.1f} us  {tf:7.1f} TFLOP/s  {k['bytes']/(k['ms']*1e-3)/1e9:8.1f} GB/s")
    eng.close()
