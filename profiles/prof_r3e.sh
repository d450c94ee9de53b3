#!/bin/bash
# round 2, pass 3e: ncu --set full of opt_step4_kernel and clv4_kernel inside the batched smoother (256 trees of 64 x 10k)
mkdir -p gpurun_out
T=256 python profiles/t_pll_batch.py > gpurun_out/r3e_plain.log 2>&1 &&
T=256 ncu --set full --clock-control none --import-source on -k regex:"opt_step4|clv4_kernel" -s 2600 -c 6 -o gpurun_out/prof_smooth_r3e python profiles/t_pll_batch.py > gpurun_out/r3e_ncu.log 2>&1
ls -la gpurun_out/prof_smooth_r3e.ncu-rep; tail -2 gpurun_out/r3e_plain.log
