#!/bin/bash
Q='python tools/quick_perf.py --only transpose --lap 4096 --st 128 --rmat 0'
for a in 0 1 0 1; do echo affine=$a; CSB200_MIRROR_AFFINE=$a $Q 2>&1 | grep "transpose\[mirror\]" | cut -c1-110; done
