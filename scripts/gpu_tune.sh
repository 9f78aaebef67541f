#!/bin/bash
mkdir -p gpurun_out
echo "== selftest"; timeout 120 ./tests/csrc/tc_selftest | tail -2
echo "== pytest tc"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider -k "tcgen05 or 3xtf32 or graph_and_eager or with_indices" 2>&1 | tail -3
for st in 0 1; do
  export DQN_STREAMS=$st
  echo "== streams=$st"; timeout 300 python bench.py --quick --steps 200 --warmup 20 2>&1 | tail -1
done
export DQN_STREAMS=1
echo "== tc report (per-kernel)"; timeout 600 python scripts/tc_report.py > gpurun_out/tc_report.log 2>&1; grep -A 34 "3xtf32-tcgen05\] eager" gpurun_out/tc_report.log; grep "vs fp64" gpurun_out/tc_report.log
