#!/bin/bash
# round 2, pass d: walk4m with the fragment table; emit order, fused chains, wide (32-byte) operand / stack / parking accesses
mkdir -p gpurun_out
python profiles/t_walk_variants.py "" variants/x_late.so variants/x_fusec.so variants/x_wide.so variants/x_wide_fusec.so variants/x_wide_late.so variants/x_wide_b2.so variants/x_b2.so > gpurun_out/r2d_variants.log 2>&1
cat gpurun_out/r2d_variants.log
PLG_LIB_PATH=$PWD/variants/x_wide_fusec.so python -m pytest tests/test_neighbourhood_gpu.py tests/test_closed_form_gpu.py -x -q 2>&1 | tail -4
