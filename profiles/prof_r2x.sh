#!/bin/bash
# round 2, pass x: opt_step4_kernel residency variants (batched smoother, 1024 trees of 64 x 10k)
for v in "" variants/opt4_mb3.so variants/opt4_mb6.so; do
  echo "== ${v:-default (5 blocks/SM)}"; PLG_LIB_PATH=$v T=1024 python profiles/t_pll_batch.py 2>&1 | head -1
done
python - <<'PY'
import sys
sys.path.insert(0, '.')
import bench
print(bench.smoothed_lnl_leg())
PY
