#!/bin/bash
run() { env "$@" timeout 300 python tools/rmat_probe.py --scale 24 --iters 5 --no-transpose --plans auto 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_median'],4), round(d['ms_best'],4))"; }
run CSB200_SPLIT_PRIO=0
run CSB200_SPLIT_PRIO=1
run CSB200_SPLIT_PRIO=0
run CSB200_SPLIT_PRIO=1
