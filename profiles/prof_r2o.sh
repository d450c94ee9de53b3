#!/bin/bash
# round 2, pass o: ncu --set full of two tbr20_step_kernel launches after the per-tile scaling decision
mkdir -p gpurun_out
python profiles/t_aa.py > gpurun_out/r2o_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tbr20_step -s 10 -c 2 -o gpurun_out/prof_tbr20_step_r2o python profiles/t_aa.py > gpurun_out/r2o_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
