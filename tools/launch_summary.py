#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel, its
share of the summed device time, launch count and mean duration.
    python tools/launch_summary.py gpurun_out/launches.csv "<command that was profiled>" """
import csv
import re
import sys
from collections import defaultdict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if not l.startswith("=="))]
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
unit = None
for r in rows[1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"cgx::", "", r[ki])
    name = re.sub(r"\(.*", "", name) if "<" not in name else re.sub(r">\(.*", ">", name)
    agg[name][0] += 1
    agg[name][1] += float(r[vi].replace(",", ""))
    unit = r[ui]
tot = sum(v[1] for v in agg.values())
print(f"ncu --metrics gpu__time_duration.sum --clock-control none  {sys.argv[2] if len(sys.argv) > 2 else ''}")
print(f"(cold-cache, serialised launches: compare SHARES)  unit: {unit}")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    per = t / n / 1e3 if unit in ("ns", "nsecond") else t / n
    print(f"{100 * t / tot:6.2f}%  n={n:5d}  avg={per:9.2f} us  {name}")
