#!/bin/bash
# ncu captures of round 2 (one GPU; never a multi-rank command): launch list of three eager steps and a full-set capture of every kernel of
# one eager step (single lane so that the launch order is the schedule's order), both of BASELINE.json configs[2] in DQN_MATH_3XTF32.
mkdir -p gpurun_out
export DQN_STREAMS=0
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python scripts/one_step.py 1 3 > gpurun_out/r02_launches.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r02_launches.log
echo "== full set, one step (second step of two)"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:tc_gemm_kernel|conv1_fwd_kernel|gather_rows|adam_kernel|sample_kernel|head_loss|colsum" -s 28 -c 28 -f -o gpurun_out/r02_step_full python scripts/one_step.py 1 2 > gpurun_out/r02_step_full.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/r02_step_full.log
ls -la gpurun_out/*.ncu-rep
