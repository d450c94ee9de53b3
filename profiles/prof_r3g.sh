#!/bin/bash
# round 2, pass 3g: the default bench line of the final code (with the full smoothed neighbourhood leg)
mkdir -p gpurun_out
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r3g.json 2> gpurun_out/bench_r3g.err ) 2>&1 | tail -3
tail -c 300 gpurun_out/bench_r3g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r3g.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'])
for k,v in d['other_workloads'].items(): print(k, round(v['ms_per_step'],3), '%.3g'%v['value'], round(v['roofline']['frac'],3), [round(x,1) for x in v['e2e']['ms_samples']])
s=d['smoothed_lnl']; print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items() if k not in ('workload','full_neighbourhood')})
print(json.dumps(s.get('full_neighbourhood'))[:900])
PY
