#!/bin/bash
# round 2, pass h: fused Newton-step kernel of the smoother, per-handle path on the device, climb semantics
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2h_tests.log
tail -5 gpurun_out/r2h_tests.log
python - <<'PY'
import sys, time, numpy as np
sys.path.insert(0, '.')
import bench
print(bench.smoothed_lnl_leg())
import os
os.environ["PLG_PLL_HANDLE_MODE"] = "host"
print("host loop:", bench.smoothed_lnl_leg(T=64)["per_tree_handle_ms"])
PY
