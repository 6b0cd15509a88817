#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "s": 1e6}.get(row["Metric Unit"], 1)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'total us':>12} {'share':>6} {'n':>5} {'avg us':>10}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:12.1f} {100 * t / tot:5.1f}% {n:5d} {t / n:10.1f}  {k[:110]}")
print(f"{tot:12.1f} total")
