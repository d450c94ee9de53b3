#!/bin/bash
# round 2, pass 3a: walk4m2_kernel (categories split over a warp pair) -- correctness on the planner tests, then
# residency variants against the shipped walk4m_kernel on the configs[1] and the target shape
mkdir -p gpurun_out
PLG_WM_SPLIT=1 timeout 900 python -m pytest tests/test_neighbourhood_gpu.py tests/test_closed_form_gpu.py -m gpu -x -q -k "planners or lnl or closed or spr" > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3a_tests.log
tail -4 gpurun_out/r3a_tests.log
echo "== shipped walk4m"; REPS=10 python profiles/t_walk_variants.py "" 2>&1 | tail -1
echo "== split variants"; PLG_WM_SPLIT=1 REPS=10 python profiles/t_walk_variants.py variants/w2_8x2.so variants/w2_6x3.so variants/w2_4x4.so variants/w2_4x5.so variants/w2_8x3.so 2>&1 | tail -5
echo "== target shape"; NT=256 NS=100000 REPS=2 python profiles/t_walk_variants.py "" 2>&1 | tail -1
PLG_WM_SPLIT=1 NT=256 NS=100000 REPS=2 python profiles/t_walk_variants.py variants/w2_8x2.so variants/w2_6x3.so variants/w2_4x5.so 2>&1 | tail -3
