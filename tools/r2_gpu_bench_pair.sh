#!/bin/bash
# usage: tools/r2_gpu_bench_pair.sh TAG [LIB]  -> short c2 + c3 bench lines under gpurun_out/
TAG=$1
if [ -n "$2" ]; then export WB200_LIB=$PWD/$2; fi
mkdir -p gpurun_out
python bench.py --steps 10 > gpurun_out/r2_bench_c2_$TAG.json 2>gpurun_out/err_$TAG.log
python bench.py --workload c3 --steps 10 --no-cpu-baseline > gpurun_out/r2_bench_c3_$TAG.json 2>>gpurun_out/err_$TAG.log
python - <<PY
import json
for w in ("c2", "c3"):
    try:
        j = json.load(open(f"gpurun_out/r2_bench_{w}_$TAG.json"))
        print(f"$TAG {w}: value {j['value']/1e6:.1f} M  warm-up {j['warmup_phase']['grad_evals_per_sec']/1e6:.1f} M  e2e {j['e2e']['value']/1e6:.1f} M  ms/step {j['ms_per_step']:.3f}  posterior {j['posterior_check']}")
    except Exception as e:
        print("$TAG", w, "failed", e)
PY
tail -2 gpurun_out/err_$TAG.log
