#!/usr/bin/env python
"""profiles/r2_chain_kernel_instr_per_eval.json from an ncu capture of `bench.py`:
executed warp instructions (all, and fp64: DFMA / DADD / DMUL / DSETP) per gradient
evaluation of the chain-kernel instances, for bench.py's `roofline_binding`.

usage: ncu_instr_per_eval.py SASS.csv RAW.csv KEY=kernel_index:evals_per_launch ... OUT.json
  SASS.csv: ncu -i X.ncu-rep --page source --csv --print-source sass
  RAW.csv : ncu -i X.ncu-rep --page raw --csv"""
import collections
import csv
import json
import re
import sys

sass_csv, raw_csv = sys.argv[1:3]
out_path = sys.argv[-1]
specs = sys.argv[3:-1]
rows = list(csv.reader(open(sass_csv)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
raw = list(csv.reader(open(raw_csv)))
rh = raw[0]
result = {}
for spec in specs:
    key, rest = spec.split("=")
    kidx, evals = rest.split(":")
    kidx, evals = int(kidx), float(evals)
    s0 = starts[kidx]
    s1 = starts[kidx + 1] if kidx + 1 < len(starts) else len(rows)
    h = rows[s0 + 1]
    ci = {n: i for i, n in enumerate(h)}
    ops = collections.Counter()
    for r in rows[s0 + 2:s1]:
        if len(r) != len(h):
            continue
        m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ci["Source"]])
        if m:
            ops[m.group(1)] += int(float(r[ci["Instructions Executed"]] or 0))
    total = sum(ops.values())
    fp64 = sum(v for k, v in ops.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
    rr = raw[2 + kidx // 1] if len(raw) > 2 + kidx else None

    def metric(name):
        try:
            v = float(rr[rh.index(name)])
            return None if v != v else v
        except Exception:
            return None

    def to_bytes(name):
        v = metric(name)
        if v is None:
            return None
        unit = raw[1][rh.index(name)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)

    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    regs = metric("launch__registers_per_thread")
    result[key] = {
        "kernel": rows[s0][1], "evals_per_launch": evals,
        "warp_instr_per_eval": total / evals, "fp64_warp_instr_per_eval": fp64 / evals,
        "top_opcodes_pct": {k: round(100 * v / total, 2) for k, v in ops.most_common(12)},
        "sm__inst_executed_pipe_fp64_pct":
            metric("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "smsp__issue_active_pct": metric("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "registers": regs,
        "resident_warps_per_sm": metric("sm__warps_active.avg.per_cycle_active"),
        "gpu_time_ms": metric("gpu__time_duration.sum"),
        "dram_bytes_per_launch": (rd + wr) if rd is not None and wr is not None else None,
        "source": sass_csv.split("/")[-1],
    }
json.dump(result, open(out_path, "w"), indent=1)
print(json.dumps(result, indent=1))
