#!/bin/bash
# round 2, pass 3p: single-launch smoother with flagged-word sums + software pipeline: tests, per-handle time,
# kernel time from an ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pll_gpu.py -x -q > gpurun_out/r3p_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r3p_tests.log
T=16 timeout 300 python profiles/t_pll_batch.py > gpurun_out/r3p_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/r3p_plain.log
T=16 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:smooth1 --csv --log-file gpurun_out/r3p_launches.csv python profiles/t_pll_batch.py > gpurun_out/r3p_ncu.log 2>&1
grep smooth1 gpurun_out/r3p_launches.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -20 | tr '\n' ' '
