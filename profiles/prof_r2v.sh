#!/bin/bash
# round 2, pass v: ncu --set full of walk4m_kernel on the headline workload (256 taxa x 100k sites, all 255,530 SPR neighbours)
mkdir -p gpurun_out
NT=256 NS=100000 REPS=1 python profiles/t_walk_variants.py "" > gpurun_out/r2v_plain.log 2>&1 &&
NT=256 NS=100000 REPS=1 ncu --set full --clock-control none --import-source off -k regex:walk4m -s 1 -c 1 -o gpurun_out/prof_walk4m_target_r2v python profiles/t_walk_variants.py "" > gpurun_out/r2v_ncu.log 2>&1
tail -2 gpurun_out/r2v_ncu.log; tail -1 gpurun_out/r2v_plain.log; ls -la gpurun_out/*r2v*
