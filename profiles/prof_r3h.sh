#!/bin/bash
# round 2, pass 3h: ncu --set full (with source) of fitch_search_walk_kernel on the configs[2] shape
mkdir -p gpurun_out
python profiles/t_fitch_kernel.py > gpurun_out/r3h_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fitch_search_walk -s 3 -c 1 -o gpurun_out/prof_fitch_sw_r3h python profiles/t_fitch_kernel.py > gpurun_out/r3h_ncu.log 2>&1
tail -1 gpurun_out/r3h_plain.log; ls -la gpurun_out/prof_fitch_sw_r3h.ncu-rep
