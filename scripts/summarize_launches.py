#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of the step)."""
import collections
import csv
import re
import sys

path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/launches.csv"
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg, tot = collections.OrderedDict(), 0.0
for row in r:
    v = float(row[vi].replace(",", ""))
    v = v / 1000 if row[ui] == "ns" else (v * 1000 if row[ui] == "ms" else v)
    name = re.sub(r"\(.*", "", row[ki])
    name = re.sub(r"^void ", "", name)[:80]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print(f"{'total us':>10s} {'share':>6s} {'n':>5s} {'avg us':>9s}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} {100 * t / tot:5.1f}% {n:5d} {t / n:9.1f}  {k}")
print(f"{tot:10.1f} total")
