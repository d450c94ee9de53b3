#!/bin/bash
# round 2, pass l: TBR-20 step kernel variants (loads up-front; 3 vs 4 blocks per SM), tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_neighbourhood_gpu.py tests/test_configs_gpu.py tests/test_lik_gpu.py -m gpu -x -q -k "tbr or c5 or protein" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2l_tests.log
tail -5 gpurun_out/r2l_tests.log
echo "== default (3 blocks/SM)"; python profiles/t_aa.py 2>&1 | tail -1
echo "== 4 blocks/SM"; PLG_LIB_PATH=variants/t20_mb4.so python profiles/t_aa.py 2>&1 | tail -1
echo "== old path"; PLG_TBR20=0 python profiles/t_aa.py 2>&1 | tail -1
