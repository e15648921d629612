#!/bin/bash
python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --timeout 120 -p no:cacheprovider -k "gaxpy" 2>&1 | tail -2
R='python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans split'
for cfg in "8 8" "4 8" "8 6" "8 4" "4 12" "4 16"; do
  set -- $cfg
  CSB200_LONG_UNR=$1 CSB200_LONG_CTAS=$2 timeout 200 $R 2>&1 | grep cs_gaxpy | sed "s/^/unr=$1 ctas=$2: /"
done > gpurun_out/r2u_long.log
cat gpurun_out/r2u_long.log | cut -c1-150
