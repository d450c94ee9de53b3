#!/bin/bash
# round 2, pass 3k: final validation -- GPU tests, smoke(), default bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r3k_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3k_tests.log
tail -3 gpurun_out/r3k_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r3k.json 2> gpurun_out/bench_r3k.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r3k.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['e2e']['value'], d['gpu_launches'])
for k,v in d['other_workloads'].items(): print(k, round(v['ms_per_step'],3), '%.3g'%v['value'], round(v['roofline']['frac'],3))
s=d['smoothed_lnl']; print({k:(round(v,3) if isinstance(v,float) else v) for k,v in s.items() if k not in ('workload','full_neighbourhood')})
f=s['full_neighbourhood']; print(f['seconds'], f['ms_per_tree'], f['roofline']['frac'])
PY
