#!/bin/bash
# round 2, final: ncu launch list of the default bench command (headline workload only), after the plain run exited 0
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r3z_plain.json 2> gpurun_out/r3z_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r3z.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/r3z_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/launches_r3z.csv")))
st = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[st]; k = h.index("Kernel Name"); v = h.index("Metric Value")
tot = collections.Counter(); n = collections.Counter()
for r in rows[st + 1:]:
    if len(r) == len(h):
        name = r[k].split("(")[0]; tot[name] += float(r[v].replace(",", "")); n[name] += 1
s = sum(tot.values())
for name, t in tot.most_common(8): print("%-40s %4d launches %12.3f ms  %5.1f %%" % (name[:40], n[name], t / 1e6, 100 * t / s))
PY
