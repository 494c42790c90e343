#!/bin/bash
# source-level ncu capture of the c3 sampling kernel (run under gpurun); brings back the
# per-instruction CSV so that the executed instruction footprint can be read off line
P=/tmp/prof; mkdir -p $P gpurun_out
B="python bench.py --no-extra-workloads --no-cpu-baseline --steps 3 --warmup 3 --workload c3"
ncu --set full --import-source on --clock-control none -k regex:walnuts_chain_kernel --launch-skip 130 -c 1 -o $P/chain_sampling_c3 $B > /dev/null 2>$P/p4.err
ncu -i $P/chain_sampling_c3.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_chain_sampling_c3.sass.csv 2>/dev/null
python tools/ncu_summary.py $P/chain_sampling_c3.ncu-rep > gpurun_out/r2_ncu_chain_sampling_c3.csv
mkdir -p $P/cubin; (cd $P/cubin && cuobjdump -xelf all $OLDPWD/walnuts_b200/csrc/build/engine.o > /dev/null)
python tools/ncu_lines.py gpurun_out/r2_chain_sampling_c3.sass.csv $P/cubin/engine.sm_100a.cubin "FunnelTargetTILi32ELi2EdEELi32ELi2ELi128ELi4ELb0Ed" 0 > gpurun_out/r2_lines_chain_sampling_c3.txt 2>&1
tail -3 $P/p4.err; ls -la gpurun_out | tail -5
