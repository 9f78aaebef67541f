#!/bin/bash
# One gpurun call for a round checkpoint: smoke + parity tests + bench (both arms) + ncu launch list + full capture of the top kernels.
mkdir -p gpurun_out
T0=$(date +%s); stamp() { echo "[t+$(( $(date +%s) - T0 ))s] $*"; }
stamp start
bash scripts/gpu_check.sh
stamp "check done"
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
stamp "reference arm done"
MODE=1 KREGEX=${KREGEX:-"tc_gemm_kernel|gather_rows_kernel|adam_kernel"} KCOUNT=${KCOUNT:-20} TAG=tc bash scripts/gpu_profile.sh
stamp "profile done"
