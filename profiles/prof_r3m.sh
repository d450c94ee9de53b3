#!/bin/bash
# round 2, pass 3m: ncu --set full (with source) of tbr20_step_kernel after the address hoist (same middle level)
mkdir -p gpurun_out
python profiles/t_aa.py > gpurun_out/r3m_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tbr20_step -s 10 -c 1 -o gpurun_out/prof_tbr20_step_r3m python profiles/t_aa.py > gpurun_out/r3m_ncu.log 2>&1
ls -la gpurun_out/prof_tbr20_step_r3m.ncu-rep
