#!/bin/bash
R='python tools/rmat_probe.py --scale 22 --iters 2 --no-gaxpy'
$R 2>&1 | tail -1 | cut -c1-160
CSB200_RS_STAGED=1 $R 2>&1 | tail -1 | sed 's/^/staged: /' | cut -c1-160
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rs_pass -s 6 -c 3 -o gpurun_out/r2x_gather -f $R > gpurun_out/r2x_ncu_g.log 2>&1; echo rc=$?
CSB200_RS_STAGED=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_rs_pass -s 6 -c 3 -o gpurun_out/r2x_staged -f $R > gpurun_out/r2x_ncu_s.log 2>&1; echo rc=$?
for n in gather staged; do ncu -i gpurun_out/r2x_$n.ncu-rep --page raw --csv > gpurun_out/r2x_$n.raw.csv; done
