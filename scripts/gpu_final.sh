#!/bin/bash
# Round checkpoint on ONE GPU: smoke, the whole GPU suite, the bench lines of every workload (both arms of the headline), the accuracy /
# per-kernel report, the stage trace, and the ncu captures of one step.  Artefacts land in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3; stamp smoke
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3; stamp tests
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/n1.err; tail -c 200 gpurun_out/n1.err; stamp bench
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r02_bench_ref.json 2> gpurun_out/ref.err; stamp reference
python bench.py --workload mlp > gpurun_out/r02_bench_mlp.json 2> gpurun_out/mlp.err; stamp mlp
python bench.py --workload drqn > gpurun_out/r02_bench_drqn.json 2> gpurun_out/drqn.err; stamp drqn
python scripts/tc_report.py > gpurun_out/r02_tc_report.txt 2>&1; stamp report
timeout 200 ./tests/csrc/tc_selftest_trace bench 20 > gpurun_out/r02_trace.log 2>&1; stamp trace
bash scripts/gpu_profile.sh; stamp profile
python - <<'PY'
import json
for f in ("n1", "ref", "mlp", "drqn"):
    try:
        d = json.loads(open(f"gpurun_out/r02_bench_{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("value"), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("e2e") or {}).get("sync_value"), (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as x:
        print(f, "FAILED", x)
PY
