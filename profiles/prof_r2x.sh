#!/bin/bash
# round 2, pass x: opt_step4_kernel with the eigensystem as kernel parameters; smoothing tests; launch list of a 256-tree batch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pll_gpu.py tests/test_configs_gpu.py -m gpu -x -q -k "LogLikelihood or smooth or climb or handle" > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2x_tests.log
tail -3 gpurun_out/r2x_tests.log
T=1024 python profiles/t_pll_batch.py 2>&1 | head -1
T=256 python profiles/t_pll_batch.py > gpurun_out/r2x_plain.log 2>&1 &&
T=256 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2x.csv python profiles/t_pll_batch.py > gpurun_out/r2x_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_r2x.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    n=r[ki].split('(')[0][:40]; tot[n]+=v; cnt[n]+=1
s=sum(tot.values())
for n,v in tot.most_common(6): print("%-42s %6d launches %10.3f ms %5.1f %%" % (n, cnt[n], v/1e6, 100*v/s))
PY
