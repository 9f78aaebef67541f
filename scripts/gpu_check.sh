#!/bin/bash
# One gpurun call: smoke, GPU parity tests (fp32 path, then tcgen05 path in its own process), numerical report, benches.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -3 gpurun_out/smoke.log
echo "== pytest gpu (fp32 path)" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider -k "not tcgen05 and not 3xtf32" > gpurun_out/pytest_gpu_fp32.log 2>&1 ; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu_fp32.log | head -40
echo "== pytest gpu (tcgen05 path)" ; timeout 900 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider -k "tcgen05 or 3xtf32" > gpurun_out/pytest_gpu_tc.log 2>&1 ; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_gpu_tc.log | head -40
echo "== tc report" ; timeout 600 python scripts/tc_report.py > gpurun_out/tc_report.log 2>&1 ; echo "tc_report rc=$?" ; head -120 gpurun_out/tc_report.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
echo "== bench" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-100} --warmup 10 --cpu-steps 10 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; cat gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
echo "== bench tc" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-100} --warmup 10 --no-cpu --math 3xtf32 > gpurun_out/bench_tc.json 2> gpurun_out/bench_tc.err ; echo "bench rc=$?" ; cat gpurun_out/bench_tc.json ; tail -5 gpurun_out/bench_tc.err
fi
