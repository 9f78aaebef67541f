"""A few eager (non-graph) batch_train! steps of config 3 for ncu captures: python scripts/one_step.py <math_mode> [steps] [N] [capture]
With `capture` the last step runs in the engine's profiling mode (single lane, one named scope per operation) between
cudaProfilerStart/Stop - run ncu with `--profile-from-start off` - and the ordered scope names go to gpurun_out/step_scopes.txt."""
import ctypes
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dqn_b200 as lib       # noqa: E402
import oracle as O           # noqa: E402
import util                  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
N = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
capture = len(sys.argv) > 4
spec = util.SPECS["c3_conv"]
cfg = lib.make_config(util.layer_descs(spec), (84, 84, 4), 6, obs_dtype="u8", batch_size=256, buffer_size=N, learning_rate=1e-4,
                      discount=0.99, seed=2, math_mode=mode, use_graph=False)
eng = lib.Engine(cfg)
eng.set_params(O.flat_params(util.make_oracle_net(spec, True, seed=1)), 0)
eng.sync_target()
eng.replay_fill_synthetic(N, 1000)
for _ in range(steps - (1 if capture else 0)):
    print(eng.train_step())
if capture:
    import torch  # noqa: F401  (its bundled libcudart is the one loaded below)
    rt = None
    for pat in ("nvidia/cuda_runtime/lib/libcudart.so*", "torch/lib/libcudart*.so*"):
        for p in glob.glob(os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), pat)):
            rt = ctypes.CDLL(p)
            break
        if rt:
            break
    eng.set_profiling(1)
    eng.sync()
    rt.cudaProfilerStart()
    print(eng.train_step())
    rt.cudaProfilerStop()
    prof = eng.get_profile()
    eng.set_profiling(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "step_scopes.txt"), "w") as f:
        for p in prof:
            f.write(f"{p['name']} {p['ms'] * 1e3:.2f}us flops={p['flops']:.0f} bytes={p['bytes']:.0f}\n")
eng.close()
