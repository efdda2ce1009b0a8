#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of total).
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h, data = rows[hdr], rows[hdr + 1:]
ki, vi, ui, gi, bi = (h.index(c) for c in ("Kernel Name", "Metric Value", "Metric Unit", "Grid Size", "Block Size"))
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1000 if r[ui] == "ns" else v * 1000 if r[ui] == "ms" else v
    name = r[ki].split("(")[0].replace("void ", "")
    a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"# {sys.argv[1]}: {len(data)} launches, {tot:.1f} us total (per-launch times are cold-cache and serialised: compare shares)")
print(f"{'us_total':>10} {'n':>5} {'us/launch':>10} {'share':>6}  grid block  kernel")
for name, (n, t, g, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {n:5d} {t / n:10.2f} {100 * t / tot:5.1f}%  {g} {b}  {name[:150]}")
