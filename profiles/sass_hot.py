#!/usr/bin/env python
"""Stall / opcode breakdown of one kernel from an ncu report captured with --import-source on.
Usage: python profiles/sass_hot.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
from collections import Counter
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr, data = rows[start], [r for r in rows[start + 1:] if len(r) == len(rows[start])]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print("samples", int(tot), "sass instructions", len(data))
for k in ["stall_long_sb", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_barrier", "stall_lg",
          "stall_not_selected", "stall_selected", "stall_dispatch", "stall_branch_resolving", "stall_no_inst"]:
    if k in ix: print("  %-24s %5.1f %%" % (k, 100 * sum(f(r, k) for r in data) / tot))
c, s = Counter(), Counter()
for r in data:
    src = r[ix["Source"]].split()
    op = (src[1] if src[0].startswith("@") else src[0]).split(".")[0]
    c[op] += f(r, "Instructions Executed"); s[op] += f(r, "# Samples")
te = sum(c.values())
print("opcode      executed  samples")
for op, n in c.most_common(18): print("  %-10s %5.1f %%  %5.1f %%" % (op, 100 * n / te, 100 * s[op] / tot))
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print("%6d %s  %-60s long %d short %d wait %d barrier %d" % (f(r, "# Samples"), r[ix["Address"]][-5:], r[ix["Source"]][:60],
          f(r, "stall_long_sb"), f(r, "stall_short_sb"), f(r, "stall_wait"), f(r, "stall_barrier")))
