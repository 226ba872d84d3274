#!/usr/bin/env python
"""Per-kernel DRAM throughput from an ncu launch list with
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv
usage: summarize_dram.py launches.csv [peak_GBps] > summary.md"""
import csv
import re
import sys
from collections import defaultdict

peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6533.8
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ki, mi, vi, ui, idi = (hdr.index(x) for x in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}
per = defaultdict(dict)
names = {}
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("gx::", "").replace("void ", "")
    name = re.sub(r"<.*", "<>", name) if name.startswith("cub::") else name
    names[r[idi]] = name
    per[r[idi]][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
agg = defaultdict(lambda: [0, 0.0, 0.0])
for i, m in per.items():
    a = agg[names[i]]
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.0)
    a[2] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
print("| kernel | launches | total us | DRAM MB | DRAM GB/s | % of measured peak |\n|---|---:|---:|---:|---:|---:|")
for k, (c, us, by) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = by / us / 1e3 if us else 0.0
    print(f"| `{k}` | {c} | {us:.1f} | {by / 1e6:.1f} | {gbs:.0f} | {100 * gbs / peak:.1f}% |")
print(f"\npeak = {peak} GB/s (MEASURED_PEAKS.json hbm_gbs); durations are ncu's (cold cache, serialised launches)")
