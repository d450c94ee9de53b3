#!/bin/bash
# round 2, pass k: 20-state TBR -- fused candidate steps + DMMA pair grid
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_neighbourhood_gpu.py tests/test_configs_gpu.py -m gpu -x -q -k "tbr or c5" > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2k_tests.log
tail -15 gpurun_out/r2k_tests.log
for mode in 1 0; do
PLG_TBR20=$mode timeout 600 python bench.py --workload aa_c5 --steps 5 --warmup 3 --no-extra --no-cpu > gpurun_out/bench_r2k_aa_c5_$mode.json 2> gpurun_out/bench_r2k_aa_c5_$mode.err
echo "mode $mode rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2k_aa_c5_$mode.json'))
print(d['ms_per_step'], d['value'], json.dumps(d['roofline'])[:1500])
PY
done
