python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1i.json 2> gpurun_out/bench_r1i.err; tail -3 gpurun_out/bench_r1i.err
python - <<'PY'
import json
b=json.load(open('gpurun_out/bench_r1i.json'))
print(b['value'], b['ms_per_step'], b['roofline']['avg_launch_ms'], b['e2e']['ms_per_step'], b['parsimony']['ms_per_step'])
PY
