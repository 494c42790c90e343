#!/bin/bash
# ncu evidence for round 2 (run under gpurun).  The .ncu-rep files stay on the box (160 MB);
# what comes back under gpurun_out/ are their summaries: key metrics (tools/ncu_summary.py),
# dynamic opcode mix (tools/ncu_opmix.py), hottest source lines (tools/ncu_lines.py), the
# per-evaluation instruction counts bench.py's roofline_binding reads, and the launch list.
P=/tmp/prof; mkdir -p $P gpurun_out
B="python bench.py --no-extra-workloads --no-cpu-baseline --steps 3 --warmup 3"
NCU="ncu --set full --import-source on --clock-control none"
$NCU -k regex:walnuts_chain_kernel -c 1 -o $P/chain_adaptive_c2 $B > /dev/null 2>$P/p1.err
$NCU -k regex:walnuts_chain_kernel --launch-skip 2 -c 1 -o $P/chain_sampling_c2 $B > /dev/null 2>$P/p2.err
$NCU -k regex:walnuts_chain_kernel --launch-skip 2 -c 1 -o $P/chain_sampling_c2_f32 $B --dtype f32 > /dev/null 2>$P/p3.err
# c3: launches 0 warm-up, 1-120 unstored sampling, 121-123 / 124-126 quota steps (W / K),
# 127-129 / 130-132 free-running steps (W / K)
$NCU -k regex:walnuts_chain_kernel --launch-skip 125 -c 1 -o $P/chain_sampling_c3 $B --workload c3 > /dev/null 2>$P/p4.err
$NCU -k regex:walnuts_chain_kernel --launch-skip 131 -c 1 -o $P/chain_free_c3 $B --workload c3 > /dev/null 2>$P/p4b.err
$NCU -k regex:"gemm_kmajor|walnuts_tick_kernel" --launch-skip 9300 -c 3 -o $P/logistic_c4 $B --workload c4 > /dev/null 2>$P/p5.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_bench_c2.csv $B > /dev/null 2>$P/p6.err
mkdir -p $P/cubin; (cd $P/cubin && cuobjdump -xelf all $OLDPWD/walnuts_b200/csrc/build/engine.o > /dev/null)
for n in chain_adaptive_c2 chain_sampling_c2 chain_sampling_c2_f32 chain_sampling_c3 chain_free_c3 logistic_c4; do
  python tools/ncu_summary.py $P/$n.ncu-rep > gpurun_out/r2_ncu_$n.csv
  ncu -i $P/$n.ncu-rep --page source --csv --print-source sass > $P/$n.sass.csv 2>/dev/null
  ncu -i $P/$n.ncu-rep --page raw --csv > $P/$n.raw.csv 2>/dev/null
  python tools/ncu_opmix.py $P/$n.sass.csv > gpurun_out/r2_opmix_$n.txt 2>&1
done
python tools/ncu_lines.py $P/chain_sampling_c2.sass.csv $P/cubin/engine.sm_100a.cubin "DiagGaussianTargetTILi128ELi4EdEELi128ELi4ELi128ELi4ELb0EdLb0" 0 > gpurun_out/r2_lines_chain_sampling_c2.txt 2>&1
python tools/ncu_lines.py $P/chain_sampling_c3.sass.csv $P/cubin/engine.sm_100a.cubin "FunnelTargetTILi32ELi2EdEELi32ELi2ELi128ELi4ELb0EdLb0" 0 > gpurun_out/r2_lines_chain_sampling_c3.txt 2>&1
(cd $P/cubin && cuobjdump -xelf all $OLDPWD/walnuts_b200/csrc/build/engine_free.o > /dev/null)
python tools/ncu_lines.py $P/chain_free_c3.sass.csv $P/cubin/engine_free.sm_100a.cubin "FunnelTargetTILi32ELi2EdEELi32ELi2ELi128ELi4ELb0EdLb1" 0 > gpurun_out/r2_lines_chain_free_c3.txt 2>&1
python tools/ncu_lines.py $P/chain_adaptive_c2.sass.csv $P/cubin/engine.sm_100a.cubin "DiagGaussianTargetTILi128ELi4EdEELi128ELi4ELi128ELi3ELb1EdLb0" 0 > gpurun_out/r2_lines_chain_adaptive_c2.txt 2>&1
# evaluations per launch from the bench itself: 10 transitions x chains x evals/transition
python - <<'PY'
import json, subprocess, sys
def evals(extra):
    out = subprocess.run([sys.executable, "bench.py", "--no-extra-workloads", "--no-cpu-baseline",
                          "--steps", "3", "--warmup", "3"] + extra, capture_output=True, text=True).stdout
    j = json.loads(out.strip().splitlines()[-1])
    return j["grad_evals_per_transition"] * j["config"]["chains_per_gpu"] * 10, j
e2, j2 = evals([])
e3, j3 = evals(["--workload", "c3"])
ef, jf = evals(["--dtype", "f32"])
ea = j2["warmup_phase"]["grad_evals_per_sec"] * j2["warmup_phase"]["ms"] * 1e-3
json.dump({"c2_sampling": e2, "c3_sampling": e3, "c3_free": e3, "c2_sampling_f32": ef, "c2_adaptive": ea},
          open("/tmp/prof/evals.json", "w"))
print(e2, e3, ef, ea)
PY
python - <<'PY'
import json, subprocess, sys
ev = json.load(open("/tmp/prof/evals.json"))
out = {}
for key, name in (("c2_sampling", "chain_sampling_c2"), ("c2_adaptive", "chain_adaptive_c2"),
                  ("c2_sampling_f32", "chain_sampling_c2_f32"), ("c3_sampling", "chain_sampling_c3"),
                  ("c3_free", "chain_free_c3")):
    subprocess.run([sys.executable, "tools/ncu_instr_per_eval.py", f"/tmp/prof/{name}.sass.csv",
                    f"/tmp/prof/{name}.raw.csv", f"{key}=0:{ev[key]}", f"/tmp/prof/{key}.json"],
                   capture_output=True)
    out.update(json.load(open(f"/tmp/prof/{key}.json")))
json.dump(out, open("gpurun_out/r2_chain_kernel_instr_per_eval.json", "w"), indent=1)
PY
cuobjdump -sass walnuts_b200/libwalnuts_b200.so | grep -E "UTC|LDTM|UTMA|Function" | grep -B1 -E "UTC|LDTM|UTMA" | awk '/Function/{f=$0} /UTC|LDTM|UTMA/{c[f" :: "$2]++} END{for(k in c) print c[k], k}' | sort -k2 > gpurun_out/r2_sass_tcgen05_tma.txt
tail -n 3 $P/p*.err | tail -30
ls -la gpurun_out/ | head -40; du -sh gpurun_out
