#!/bin/bash
mkdir -p gpurun_out
T=16 python profiles/t_pll_batch.py > /dev/null 2>&1 || exit 1
for cap in 40 80 148 296; do
PLG_S1_MAX_BLOCKS=$cap T=16 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:smooth1 -c 4 --csv --log-file gpurun_out/l.csv python profiles/t_pll_batch.py > /dev/null 2>&1
echo "cap $cap: $(grep smooth1 gpurun_out/l.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -2 | tr '\n' ' ') ns"
done
