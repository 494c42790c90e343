#!/bin/bash
# ncu summary of the lock-step tick's kernels at c4 (run under gpurun)
P=/tmp/prof; mkdir -p $P gpurun_out
B="python bench.py --no-extra-workloads --no-cpu-baseline --steps 3 --warmup 3 --workload c4"
ncu --set full --import-source on --clock-control none -k regex:"gemm_kmajor|walnuts_tick_kernel" --launch-skip 9300 -c 3 -o $P/logistic_c4 $B > /dev/null 2>$P/p5.err
python tools/ncu_summary.py $P/logistic_c4.ncu-rep > gpurun_out/r2_ncu_logistic_c4.csv
ncu -i $P/logistic_c4.ncu-rep --page source --csv --print-source sass > $P/logistic_c4.sass.csv 2>/dev/null
python tools/ncu_opmix.py $P/logistic_c4.sass.csv > gpurun_out/r2_opmix_logistic_c4.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 9000 -c 600 --csv --log-file gpurun_out/r2_launches_c4.csv $B > /dev/null 2>$P/p6.err
cuobjdump -sass walnuts_b200/libwalnuts_b200.so | grep -E "UTC|LDTM|UTMA|Function" | grep -B1 -E "UTC|LDTM|UTMA" | awk '/Function/{f=$0} /UTC|LDTM|UTMA/{c[f" :: "$2]++} END{for(k in c) print c[k], k}' | sort -k2 > gpurun_out/r2_sass_tcgen05_tma.txt
tail -2 $P/p5.err
