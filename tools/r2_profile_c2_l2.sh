#!/bin/bash
# one ncu --set full capture of the c2 sampling kernel; summary incl. the L2 / L1 throughput
P=/tmp/prof; mkdir -p $P gpurun_out
B="python bench.py --no-extra-workloads --no-cpu-baseline --steps 3 --warmup 3"
ncu --set full --clock-control none -k regex:walnuts_chain_kernel --launch-skip 2 -c 1 -o $P/c2l2 $B > /dev/null 2>$P/e.err
python tools/ncu_summary.py $P/c2l2.ncu-rep > gpurun_out/r2n_ncu_c2_l2.csv
ncu -i $P/c2l2.ncu-rep --page raw --csv | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h,u,d=rows[0],rows[1],rows[2]
for k,uu,v in zip(h,u,d):
    if any(t in k for t in ('lts__','l1tex__t_','throughput','xbar')): print(k,uu,v)
" > gpurun_out/r2n_ncu_c2_l2_all.txt
tail -3 $P/e.err
