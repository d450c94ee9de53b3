#!/bin/bash
# round 2, pass 3n: staged 20-state update kernel: tests, protein smoother and TBR / SPR timing against the unstaged one
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_lik_gpu.py tests/test_pll_gpu.py tests/test_neighbourhood_gpu.py tests/test_configs_gpu.py tests/test_edge_gpu.py -m gpu -x -q 2>&1 | tail -3
for st in 1 0; do
  echo "== PLG_A20_STAGED=$st"
  PLG_A20_STAGED=$st T=256 python profiles/t_pll_batch_aa.py 2>&1 | head -1
  PLG_A20_STAGED=$st python profiles/t_aa.py 2>&1 | tail -2 | cut -c1-210
done
