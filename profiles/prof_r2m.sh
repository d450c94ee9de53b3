#!/bin/bash
# round 2, pass m: TBR-20 step kernel, tile loop unrolled 8 / 2 / 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_neighbourhood_gpu.py tests/test_configs_gpu.py -m gpu -x -q -k "tbr or c5" > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2m_tests.log
tail -5 gpurun_out/r2m_tests.log
echo "== unroll 8"; python profiles/t_aa.py 2>&1 | tail -1
echo "== unroll 2"; PLG_LIB_PATH=variants/t20_u2.so python profiles/t_aa.py 2>&1 | tail -1
echo "== unroll 1"; PLG_LIB_PATH=variants/t20_u1.so python profiles/t_aa.py 2>&1 | tail -1
