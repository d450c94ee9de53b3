#!/bin/bash
# round 2, pass 3j: ncu --set full (with source) of smooth1_kernel (one 64-taxon x 10k tree)
mkdir -p gpurun_out
T=16 python profiles/t_pll_batch.py > gpurun_out/r3j_plain.log 2>&1 &&
T=16 ncu --set full --clock-control none --import-source on -k regex:smooth1 -s 2 -c 1 -o gpurun_out/prof_smooth1_r3j python profiles/t_pll_batch.py > gpurun_out/r3j_ncu.log 2>&1
ls -la gpurun_out/prof_smooth1_r3j.ncu-rep
