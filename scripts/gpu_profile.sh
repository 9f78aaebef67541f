#!/bin/bash
# ncu captures (one GPU): launch list of a few eager steps + full-set capture of selected kernels.
mkdir -p gpurun_out
MODE=${MODE:-0}
echo "== launch list (mode $MODE)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_m${MODE}.csv python scripts/one_step.py $MODE 3 > gpurun_out/launches_m${MODE}.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/launches_m${MODE}.log
echo "== full capture: ${KREGEX:-igemm_kernel}"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX:-igemm_kernel}" -s ${KSKIP:-0} -c ${KCOUNT:-4} -f -o gpurun_out/prof_${TAG:-k}_m${MODE} python scripts/one_step.py $MODE 1 > gpurun_out/prof_${TAG:-k}_m${MODE}.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/prof_${TAG:-k}_m${MODE}.log
ls -la gpurun_out | tail -8
