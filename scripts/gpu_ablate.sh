cp deepqlearning.jl_b200/libdqn_b200.so /tmp/lib_orig.so
for f in gpurun_in_lib_*.so; do
  cp $f deepqlearning.jl_b200/libdqn_b200.so
  echo "=== $f"; timeout 300 python scripts/tc_report.py 2>&1 | grep -A 30 "3xtf32-tcgen05\] eager" | grep -E "total|conv|dense1" | awk '{printf "%s %s | ", $1, $2} END {print ""}'
  timeout 300 python bench.py --quick --steps 200 --warmup 20 2>&1 | tail -1 | cut -c1-90
done
cp /tmp/lib_orig.so deepqlearning.jl_b200/libdqn_b200.so
