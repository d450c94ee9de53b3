#!/bin/bash
# round 2, pass i: full GPU test pass (lock support, C2/C5 parity widened), smoother regression fix, every bench workload
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2i_tests.log
tail -5 gpurun_out/r2i_tests.log
python - <<'PY'
import sys
sys.path.insert(0, '.')
import bench
print(bench.smoothed_lnl_leg())
PY
for w in aa_c5 trees_c4 fitch_target lnl_c2; do
  python bench.py --workload $w --steps 5 --warmup 3 --no-extra > gpurun_out/bench_r2i_$w.json 2> gpurun_out/bench_r2i_$w.err
  echo "$w rc=$?"; head -c 600 gpurun_out/bench_r2i_$w.json; echo
done
