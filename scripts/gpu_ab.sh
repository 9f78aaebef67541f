for m in 1 0; do export DQN_MERGE_FWD=$m; echo "== merge=$m"; timeout 300 python bench.py --quick --steps 300 --warmup 30 2>&1 | tail -1; done
export DQN_MERGE_FWD=1
timeout 600 python scripts/tc_report.py > gpurun_out/tc_report.log 2>&1; grep -A 34 "3xtf32-tcgen05\] eager" gpurun_out/tc_report.log | grep -E "fwd|reduce|total"; grep "vs fp64" gpurun_out/tc_report.log
