#!/bin/bash
run() { env "$@" timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-gaxpy 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_median'],4), round(d['ms_best'],4))"; }
run CSB200_RS_LB=1
run CSB200_RS_LB=8
run CSB200_RS_LB=1
run CSB200_RS_LB=4
