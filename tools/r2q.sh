#!/bin/bash
R='python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans split'
for cfg in "64 16 1024" "64 24 1024" "64 32 1024" "96 16 1024" "64 16 512" "128 32 1024" "96 24 1024"; do
  set -- $cfg
  CSB200_SPLIT_LONG=$1 CSB200_SPLIT_MID=$2 CSB200_SPLIT_CHUNK=$3 timeout 200 $R 2>&1 | grep cs_gaxpy | sed "s/^/long=$1 mid=$2 chunk=$3: /"
done > gpurun_out/r2q_split.log
cat gpurun_out/r2q_split.log | cut -c1-160
