#!/bin/bash
# round 2, pass c: walk4m variants that take load off the L1 data pipe (fragment table, DMMA quad sums, fused chains)
mkdir -p gpurun_out
python profiles/t_walk_variants.py "" variants/w_f256.so variants/w_qsum.so variants/w_fusec.so variants/w_fq.so variants/w_ff.so variants/w_all.so variants/w_all_b2.so > gpurun_out/r2c_variants.log 2>&1
cat gpurun_out/r2c_variants.log
PLG_LIB_PATH=$PWD/variants/w_all.so python -m pytest tests/test_neighbourhood_gpu.py tests/test_closed_form_gpu.py -x -q 2>&1 | tail -4
