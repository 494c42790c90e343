#!/usr/bin/env python
"""Dynamic instruction mix of a kernel from `ncu -i X.ncu-rep --page source --csv
--print-source sass`: executed warp instructions and stall samples per opcode, per kernel."""
import csv
import collections
import re
import sys

path = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 else None
rows = list(csv.reader(open(path)))
kernels = []
cur = None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = dict(name=r[1], header=None, rows=[])
        kernels.append(cur)
    elif cur is not None and cur["header"] is None and r and r[0] == "Address":
        cur["header"] = r
    elif cur is not None and cur["header"] is not None and r and r[0].startswith("0x") is False and len(r) == len(cur["header"]):
        cur["rows"].append(r)
    elif cur is not None and cur["header"] is not None and r:
        cur["rows"].append(r)
for k in kernels:
    if which and which not in k["name"]:
        continue
    h = k["header"]
    ci = {n: i for i, n in enumerate(h)}
    ex = collections.Counter()
    st = collections.Counter()
    tot = 0
    tots = 0
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    stall_tot = collections.Counter()
    for r in k["rows"]:
        if len(r) < len(h):
            continue
        src = r[ci["Source"]]
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
        if not m:
            continue
        op = m.group(1)
        try:
            n = int(float(r[ci["Instructions Executed"]] or 0))
            s = int(float(r[ci["# Samples"]] or 0))
        except ValueError:
            continue
        ex[op] += n
        st[op] += s
        tot += n
        tots += s
        for c in stall_cols:
            try:
                stall_tot[c] += int(float(r[ci[c]] or 0))
            except ValueError:
                pass
    print("==", k["name"][:110])
    print(f"   executed warp instructions {tot:,}; samples {tots:,}")
    print("   op        exec%   samples%")
    for op, n in ex.most_common(22):
        print(f"   {op:9s} {100*n/max(tot,1):6.2f}  {100*st[op]/max(tots,1):6.2f}")
    print("   stalls:", ", ".join(f"{c[6:]} {100*v/max(tots,1):.1f}%" for c, v in stall_tot.most_common(10)))
