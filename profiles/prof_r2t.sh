#!/bin/bash
# round 2, pass t: per-handle smoothing in one cooperative launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pll_gpu.py tests/test_dropin_gpu.py tests/test_configs_gpu.py -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2t_tests.log
tail -25 gpurun_out/r2t_tests.log
timeout 300 python - <<'PY'
import sys
sys.path.insert(0, '.')
import bench, os
print(bench.smoothed_lnl_leg(T=64))
os.environ["PLG_PLL_HANDLE_MODE"] = "batch"
print("batch-of-one:", bench.smoothed_lnl_leg(T=64)["per_tree_handle_ms"])
PY
