#!/bin/bash
# round 2, pass 3l: batched smoother on protein data: tests, wall time and ncu launch list
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pll_gpu.py -m gpu -x -q 2>&1 | tail -3
T=256 python profiles/t_pll_batch_aa.py 2>&1 | tail -2
T=128 python profiles/t_pll_batch_aa.py > gpurun_out/r3l_plain.log 2>&1 &&
T=128 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r3l.csv python profiles/t_pll_batch_aa.py > gpurun_out/r3l_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/launches_r3l.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
tot=collections.Counter(); cnt=collections.Counter()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    n=r[ki].split('(')[0][:40]; tot[n]+=v; cnt[n]+=1
s=sum(tot.values())
for n,v in tot.most_common(6): print("%-42s %6d launches %10.3f ms %5.1f %%" % (n, cnt[n], v/1e6, 100*v/s))
PY
