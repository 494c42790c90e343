#!/usr/bin/env python
"""Static SASS size per source line of one kernel (what competes for the 32 KB L1.5
instruction cache).  usage: sass_lines.py <cubin> <mangled kernel name> [top]
(cubins: cuobjdump -xelf all libwalnuts_b200.so)"""
import collections
import re
import subprocess
import sys

cubin, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
syms = subprocess.run(["readelf", "-sW", cubin], capture_output=True, text=True).stdout
idx = next(l.split()[0].rstrip(":") for l in syms.splitlines() if l.endswith(" " + kernel))
text = subprocess.run(["nvdisasm", "-g", "-fun", idx, cubin], capture_output=True,
                      text=True).stdout
cur, inside = None, False
cnt = collections.Counter()
ops = collections.Counter()
for l in text.splitlines():
    if l.startswith("//--------------------- .text."):
        inside = kernel in l
        continue
    if l.startswith("//--------------------- "):
        inside = False
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
    if m:
        cnt[cur] += 1
        ops[m.group(1)] += 1
tot = sum(cnt.values())
print("total instructions", tot, "=", tot * 16 // 1024, "KB")
print("ops:", ", ".join(f"{k} {v}" for k, v in ops.most_common(16)))
for k, v in cnt.most_common(top):
    print(f"{v:6d} {100 * v / tot:5.1f}%  {k}")
