#!/bin/bash
# round 2, pass x: opt_step4_kernel variants (blocks per SM x tiles per block), batched smoother, 1024 trees of 64 x 10k
for v in variants/o4_mb3_t1.so variants/o4_mb3_t2.so variants/o4_mb4_t2.so variants/o4_mb4_t4.so; do
  echo "== $v"; PLG_LIB_PATH=$v T=1024 python profiles/t_pll_batch.py 2>&1 | head -1
done
