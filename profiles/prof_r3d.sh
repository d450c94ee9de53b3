#!/bin/bash
# round 2, pass 3d: walk4m with the next visit's far-side operands staged by cp.async, against the same build without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_neighbourhood_gpu.py tests/test_configs_gpu.py tests/test_closed_form_gpu.py -m gpu -x -q > gpurun_out/r3d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3d_tests.log
tail -3 gpurun_out/r3d_tests.log
REPS=10 python profiles/t_walk_variants.py "" variants/wm_nostage.so 2>&1 | tail -2
NT=256 NS=100000 REPS=2 python profiles/t_walk_variants.py "" variants/wm_nostage.so 2>&1 | tail -2
