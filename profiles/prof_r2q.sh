#!/bin/bash
# round 2, pass q: ncu --set full of the current walk4m_kernel (C2 shape)
mkdir -p gpurun_out
REPS=2 python profiles/t_walk_variants.py "" > gpurun_out/r2q_plain.log 2>&1 &&
REPS=2 ncu --set full --clock-control none --import-source on -k regex:walk4m -s 2 -c 1 -o gpurun_out/prof_walk4m_r2q python profiles/t_walk_variants.py "" > gpurun_out/r2q_ncu.log 2>&1
tail -3 gpurun_out/r2q_ncu.log; cat gpurun_out/r2q_plain.log | tail -3
