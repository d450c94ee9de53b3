#!/bin/bash
# round 2, pass z: device time of smooth1_kernel (one 64-taxon x 10k tree per launch) from an ncu launch list; its tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pll_gpu.py -m gpu -x -q 2>&1 | tail -3
T=16 python profiles/t_pll_batch.py > gpurun_out/r2z_plain.log 2>&1 || exit 1
tail -1 gpurun_out/r2z_plain.log
T=16 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:smooth1 -c 6 --csv --log-file gpurun_out/l.csv python profiles/t_pll_batch.py > /dev/null 2>&1
echo "smooth1_kernel: $(grep smooth1 gpurun_out/l.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -3 | tr '\n' ' ') ns"
