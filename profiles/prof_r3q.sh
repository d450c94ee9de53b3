#!/bin/bash
# round 2, pass 3q: ncu --set full (with source) of the pipelined smooth1_kernel (one 64-taxon x 10k tree)
mkdir -p gpurun_out
T=16 timeout 900 ncu --set full --clock-control none --import-source on -k regex:smooth1 -s 2 -c 1 -o gpurun_out/prof_smooth1_r3q python profiles/t_pll_batch.py > gpurun_out/r3q_ncu.log 2>&1
ls -la gpurun_out/prof_smooth1_r3q.ncu-rep
