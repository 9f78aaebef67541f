#!/bin/bash
# quick tuning loop: selftest, tcgen05 parity tests, device-resident bench, per-kernel report
mkdir -p gpurun_out
echo "== selftest"; timeout 300 ./tests/csrc/tc_selftest > gpurun_out/selftest.log 2>&1; echo "rc=$?"; grep -E "FAIL|SELFTEST|error" gpurun_out/selftest.log | head
for f in tests/csrc/tc_selftest tests/csrc/tc_selftest_*; do [ -x $f ] && { echo "== $f bench"; timeout 120 $f bench 10 2>&1 | sed -e "s/algorithmic.*grid/grid/" | tail -12; }; done
if [ -z "$SKIP_PYTEST" ]; then
echo "== pytest tc"; timeout 900 python -m pytest tests -m gpu -q --timeout 600 --no-header -p no:cacheprovider -x -k "${PYTEST_K:-tcgen05 or 3xtf32 or graph_and_eager or with_indices}" 2>&1 | tail -5
fi
for tma in ${TMAS:-1 0}; do
  export DQN_TC_TMA=$tma
  echo "== DQN_TC_TMA=$tma"; timeout 300 python bench.py --quick --steps 200 --warmup 20 2>&1 | tail -1
done
export DQN_TC_TMA=1
echo "== tc report (per-kernel)"; timeout 600 python scripts/tc_report.py > gpurun_out/tc_report.log 2>&1; grep -A 36 "3xtf32-tcgen05\] eager" gpurun_out/tc_report.log; grep "vs fp64" gpurun_out/tc_report.log
