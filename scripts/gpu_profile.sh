#!/bin/bash
# ncu captures of round 2 (one GPU; never a multi-rank command), BASELINE.json configs[2] in DQN_MATH_3XTF32:
#   1. launch list (gpu__time_duration only) of exactly one eager step in the engine's profiling mode (single lane: launch order = schedule order)
#   2. --set full capture of every kernel of one such step
# The step is bracketed by cudaProfilerStart/Stop (scripts/one_step.py capture), so nothing else is in the reports.
mkdir -p gpurun_out
echo "== launch list"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python scripts/one_step.py 1 3 65536 capture > gpurun_out/r02_launches.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r02_launches.log; cp gpurun_out/step_scopes.txt gpurun_out/r02_step_scopes.txt
echo "== full set, one step"
timeout 1500 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r02_step_full python scripts/one_step.py 1 3 65536 capture > gpurun_out/r02_step_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r02_step_full.log
ls -la gpurun_out/*.ncu-rep
