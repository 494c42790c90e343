#!/usr/bin/env python
"""Tabulate an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel."""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[h]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot = defaultdict(lambda: [0, 0.0])
for r in rows[h + 1:]:
    if len(r) > vi:
        n = r[ki].split("(")[0][-80:]
        tot[n][0] += 1
        tot[n][1] += float(r[vi])
s = sum(v[1] for v in tot.values())
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{t / 1e6:10.3f} ms {100 * t / s:6.2f}%  x{c:<4d} {n}")
