#!/bin/bash
R='python tools/rmat_probe.py --scale 24 --iters 2 --no-transpose --plans auto'
i=0
for cfg in "CSB200_HOT_K=0 CSB200_XLD=0" "CSB200_HOT_K=8192 CSB200_XLD=1" "CSB200_HOT_K=8192 CSB200_XLD=0" "CSB200_HOT_K=16384 CSB200_XLD=0" "CSB200_HOT_K=4096 CSB200_XLD=0"; do
 i=$((i+1))
 env $cfg timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_spmv_\|k_long --csv --log-file gpurun_out/r2y_$i.csv $R > /dev/null 2>&1
 echo "$cfg"; python tools/ncu_summary.py launches gpurun_out/r2y_$i.csv | grep "k_" | cut -c1-80
done
