#!/bin/bash
# round 2, pass 3b: full GPU test pass, smoke(), and the default bench line + reference arm of the final code
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r3b_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3b_tests.log
tail -4 gpurun_out/r3b_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r3b.json 2> gpurun_out/bench_r3b.err ) 2>&1 | tail -3
( time PLG_BENCH_REF_SECONDS=30 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r3b_ref.json 2> gpurun_out/bench_r3b_ref.err ) 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r3b.json'))
print(d['ms_per_step'], d['value'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value'])
for k,v in d['other_workloads'].items(): print(k, round(v['ms_per_step'],3), '%.3g'%v['value'], round(v['roofline']['frac'],3))
print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d['smoothed_lnl'].items() if k!='workload'})
PY
