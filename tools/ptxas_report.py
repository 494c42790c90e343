#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: one line per kernel (registers, spills, stack, smem)."""
import re, sys, subprocess
txt = sys.stdin.read()
cur = None
rows = []
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = dict(name=m.group(1), regs=0, spill_st=0, spill_ld=0, stack=0, smem=0); rows.append(cur); continue
    if cur is None: continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and not cur.get("seen_stack"):
        cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups()); cur["seen_stack"] = True
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
        m2 = re.search(r"(\d+) bytes smem", line)
        if m2: cur["smem"] = int(m2.group(1))
try:
    names = subprocess.run(["c++filt"] + [r["name"] for r in rows], capture_output=True, text=True).stdout.splitlines()
except Exception:
    names = [r["name"] for r in rows]
for r, n in zip(rows, names):
    n = re.sub(r"\(.*\)$", "", n).replace("wb200::", "").replace("void ", "")
    print(f"{r['regs']:4d} regs  stack {r['stack']:5d}  spill st/ld {r['spill_st']:5d}/{r['spill_ld']:5d}  smem {r['smem']:6d}  {n}")
