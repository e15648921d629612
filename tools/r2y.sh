#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q --timeout 200 -p no:cacheprovider -k "gaxpy" 2>&1 | tail -2
run() { env "$@" timeout 300 python tools/rmat_probe.py --scale 24 --iters 5 --no-transpose --plans auto 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_median'],4), round(d['ms_best'],4))"; }
run CSB200_SHORT1=1
run CSB200_SHORT1=0
run CSB200_SPLIT_MID=32
run CSB200_SPLIT_MID=8
run CSB200_SPLIT_MID=32 CSB200_SPLIT_LONG=128
run CSB200_SPLIT_LONG=128
R='python tools/rmat_probe.py --scale 24 --iters 2 --no-transpose --plans auto'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_spmv_\|k_long --csv --log-file gpurun_out/r2y_6.csv $R > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2y_6.csv | grep "k_" | cut -c1-80
