#!/usr/bin/env python
"""Samples per SOURCE LINE of one kernel: joins an ncu report (captured with --import-source on) with nvdisasm -g
line markers of the same build.  Usage: python profiles/line_hot.py report.ncu-rep file.cubin kernel_substring [top_n]"""
import csv, re, subprocess, sys
from collections import Counter
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
line_of, cur, on = {}, None, False
for l in sass:
    if l.startswith(".text."):
        on = kern in l
        continue
    if not on: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
    if m: line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(out.splitlines()))
st = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
h = rows[st]; ix = {k: i for i, k in enumerate(h)}
data = [r for r in rows[st + 1:] if len(r) == len(h)]
base = int(data[0][ix["Address"]], 16)
samp, execd = Counter(), Counter()
for r in data:
    off = int(r[ix["Address"]], 16) - base
    k = line_of.get(off)
    samp[k] += float(r[ix["# Samples"]] or 0); execd[k] += float(r[ix["Instructions Executed"]] or 0)
tot = sum(samp.values()); te = sum(execd.values())
print("samples %d, warp instructions %d" % (tot, te))
src = {}
for k, n in samp.most_common(top):
    if k is None: print("%6.2f %%  %6.2f %% exec  ?" % (100 * n / tot, 100 * execd[k] / te)); continue
    f, ln = k
    if f not in src:
        try: src[f] = open("pylogeny_b200/csrc/" + f).read().splitlines()
        except Exception: src[f] = []
    text = src[f][ln - 1].strip()[:90] if ln - 1 < len(src[f]) else ""
    print("%6.2f %%  %6.2f %% exec  %s:%d  %s" % (100 * n / tot, 100 * execd[k] / te, f, ln, text))
