#!/bin/bash
# round 2, pass f: ncu --set full of the current walk4m_kernel (pair fragment table, no log, REDUX exponents)
mkdir -p gpurun_out
REPS=2 python profiles/t_walk_variants.py "" > gpurun_out/r2f_plain.log 2>&1 &&
REPS=2 ncu --set full --clock-control none --import-source on -k regex:walk4m -s 2 -c 1 -o gpurun_out/prof_walk4m_r2f python profiles/t_walk_variants.py "" > gpurun_out/r2f_ncu.log 2>&1
tail -3 gpurun_out/r2f_ncu.log
NT=256 NS=100000 REPS=2 python profiles/t_walk_variants.py variants/r1.so "" 2>&1 | tail -2
