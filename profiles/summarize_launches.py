#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
usage: summarize_launches.py launches.csv > summary.md"""
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("gx::", "")
    name = re.sub(r"<.*", "<>", name) if name.startswith(("cub::", "void cub::")) else name
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", "")) / 1e3
total = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {c} | {us:.1f} | {100 * us / total:.1f}% | {us / c:.1f} |")
print(f"\ntotal {total / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches (cold-cache, serialised: compare shares, not absolutes)")
