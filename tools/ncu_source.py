#!/usr/bin/env python
"""Per-source-line instruction share, sample share and top stall reasons from
`ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > file.csv`."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
h = rows[hi[0]]
idx = {n: i for i, n in enumerate(h)}
per = collections.OrderedDict()
end = hi[1] if len(hi) > 1 else len(rows)
for r in rows[hi[0] + 1:end]:
    if len(r) < len(h):
        continue
    try:
        ln = int(r[0])
    except ValueError:
        continue
    try:  # source lines with embedded quotes (inline asm) come out of ncu's CSV mis-split
        n_inst = float(r[idx["Instructions Executed"]].replace("-", "0") or 0)
        n_samp = float(r[idx["# Samples"]].replace("-", "0") or 0)
    except ValueError:
        continue
    d = per.setdefault(ln, [r[1], 0, 0, collections.Counter()])
    d[1] += n_inst
    d[2] += n_samp
    for k in h:
        if k.startswith("stall_") and "Not" not in k and r[idx[k]]:
            try:
                d[3][k] += float(r[idx[k]])
            except ValueError:
                pass
ti = sum(d[1] for d in per.values())
ts = sum(d[2] for d in per.values())
print("warp instructions", ti, "samples", ts)
for ln, d in sorted(per.items()):
    if d[1] / ti > thr or d[2] / ts > thr:
        top = ", ".join(f"{k[6:]}:{v / ts * 100:.1f}" for k, v in d[3].most_common(3))
        print(f"{ln:4d} inst {100 * d[1] / ti:5.1f}% samp {100 * d[2] / ts:5.1f}%  [{top}]  {d[0].strip()[:100]}")
