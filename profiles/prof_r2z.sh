#!/bin/bash
# round 2, pass z: device time of smooth1_kernel (one 64-taxon x 10k tree per launch) from an ncu launch list
mkdir -p gpurun_out
T=64 python profiles/t_pll_batch.py > gpurun_out/r2z_plain.log 2>&1 &&
T=64 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:smooth1 -c 8 --csv --log-file gpurun_out/launches_r2z.csv python profiles/t_pll_batch.py > gpurun_out/r2z_ncu.log 2>&1
grep -E "smooth1" gpurun_out/launches_r2z.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | head -20
tail -2 gpurun_out/r2z_plain.log
PLG_TIMING=1 T=8 python profiles/t_pll_batch.py 2>&1 | grep "\[plg\]" | tail -12
