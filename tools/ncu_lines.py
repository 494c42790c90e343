#!/usr/bin/env python
"""Join `ncu --page source --csv --print-source sass` (executed counts per SASS instruction)
with `nvdisasm -g -c` line info of the same cubin: executed warp instructions and stall
samples per CUDA source line.
usage: ncu_lines.py SASS.csv CUBIN MANGLED_SUBSTRING [KERNEL_INDEX]"""
import collections
import csv
import re
import subprocess
import sys

sass_csv, cubin, key = sys.argv[1:4]
kidx = int(sys.argv[4]) if len(sys.argv) > 4 else -1
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# per function: list of (source line) in instruction order
lines_for = []
cur = None
loc = None
infn = False
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+),", ln)
    if m:
        infn = key in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        loc = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines_for.append((int(m.group(1), 16), loc, m.group(2)))
rows = list(csv.reader(open(sass_csv)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
s0 = starts[kidx]
s1 = starts[starts.index(s0) + 1] if starts.index(s0) + 1 < len(starts) else len(rows)
h = rows[s0 + 1]
ci = {n: i for i, n in enumerate(h)}
per = collections.Counter()
smp = collections.Counter()
ops = collections.defaultdict(collections.Counter)
body = [r for r in rows[s0 + 2:s1] if len(r) == len(h)]
if len(body) != len(lines_for):
    print(f"warning: {len(body)} profiled instructions vs {len(lines_for)} disassembled", file=sys.stderr)
tot = tots = 0
for r, (off, loc, text) in zip(body, lines_for):
    n = int(float(r[ci["Instructions Executed"]] or 0))
    s = int(float(r[ci["# Samples"]] or 0))
    per[loc] += n
    smp[loc] += s
    op = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", text)
    ops[loc][op.group(1) if op else "?"] += n
    tot += n
    tots += s
print(f"kernel: {rows[s0][1][:120]}\ntotal executed {tot:,}, samples {tots:,}")
src_cache = {}
for loc, n in per.most_common(45):
    f, l = loc if loc else ("?", 0)
    top = ", ".join(f"{o}:{100*c/max(n,1):.0f}%" for o, c in ops[loc].most_common(4))
    print(f"{100*n/tot:5.2f}% exec {100*smp[loc]/max(tots,1):5.2f}% stall  {f}:{l:<4} [{top}]")
