#!/bin/bash
R='python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans split'
for cfg in "4 8" "8 8" "4 6" "4 12" "8 6" "8 4"; do
  set -- $cfg
  CSB200_LONG_UNR=$1 CSB200_LONG_CTAS=$2 timeout 200 $R 2>&1 | grep cs_gaxpy | sed "s/^/unr=$1 ctas=$2: /"
done > gpurun_out/r2r_long.log
cat gpurun_out/r2r_long.log | cut -c1-150
