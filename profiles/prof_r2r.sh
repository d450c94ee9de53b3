#!/bin/bash
# round 2, pass r: walk4m with pi folded into the pruned subtree's vector; the lnL planner tests; default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_neighbourhood_gpu.py tests/test_configs_gpu.py tests/test_closed_form_gpu.py -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2r_tests.log
tail -4 gpurun_out/r2r_tests.log
REPS=10 python profiles/t_walk_variants.py "" 2>&1 | tail -1
NT=256 NS=100000 REPS=2 python profiles/t_walk_variants.py "" 2>&1 | tail -1
