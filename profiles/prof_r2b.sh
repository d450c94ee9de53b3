#!/bin/bash
# round 2, pass b: walk4m variants (dual emit, partial in shared memory, cp.async operand staging)
mkdir -p gpurun_out
python profiles/t_walk_variants.py "" variants/v_dual.so variants/v_ak.so variants/v_ak_dual.so variants/v_st_b2.so variants/v_st_b2_dual.so variants/v_st.so variants/v_st_dual.so variants/v_ak_st_b2.so > gpurun_out/r2b_variants.log 2>&1
cat gpurun_out/r2b_variants.log
python -m pytest tests/test_neighbourhood_gpu.py tests/test_closed_form_gpu.py tests/test_treewalk_gpu.py tests/test_pll_gpu.py -x -q 2>&1 | tail -4
