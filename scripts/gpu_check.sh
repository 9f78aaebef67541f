#!/bin/bash
# One gpurun call: smoke, GPU parity tests, short bench.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1 ; echo "smoke rc=$?" ; tail -5 gpurun_out/smoke.log
echo "== pytest gpu" ; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1 ; echo "pytest rc=$?" ; tail -40 gpurun_out/pytest_gpu.log
echo "== bench" ; timeout 900 python bench.py --steps ${BENCH_STEPS:-100} --warmup 10 --cpu-steps 10 > gpurun_out/bench.json 2> gpurun_out/bench.err ; echo "bench rc=$?" ; cat gpurun_out/bench.json ; tail -5 gpurun_out/bench.err
