#!/bin/bash
# round 2, pass e: walk4m with the pair fragment table, REDUX exponents; re-fetched fragment, DMMA quad sums, block shapes
mkdir -p gpurun_out
python profiles/t_walk_variants.py "" variants/y_refrag.so variants/y_fusec.so variants/y_qd.so variants/y_qd_fusec.so variants/y_w8.so variants/y_w2.so > gpurun_out/r2e_variants.log 2>&1
cat gpurun_out/r2e_variants.log
python -m pytest tests/test_neighbourhood_gpu.py tests/test_closed_form_gpu.py tests/test_configs_gpu.py -x -q 2>&1 | tail -4
