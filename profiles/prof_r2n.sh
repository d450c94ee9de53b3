#!/bin/bash
# round 2, pass n: ncu --set full of two tbr20_step_kernel launches (middle levels of the C5 TBR neighbourhood) and of pairs20_kernel
mkdir -p gpurun_out
python profiles/t_aa.py > gpurun_out/r2n_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tbr20_step -s 10 -c 2 -o gpurun_out/prof_tbr20_step_r2n python profiles/t_aa.py > gpurun_out/r2n_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pairs20 -s 1 -c 1 -o gpurun_out/prof_pairs20_r2n python profiles/t_aa.py > gpurun_out/r2n_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep
