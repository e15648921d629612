#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py tests/test_gpu_fuzz.py -x -q --timeout 200 -p no:cacheprovider -k "transpose or compress or symperm" 2>&1 | tail -2
run() { env "$@" timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-gaxpy 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_median'],4), round(d['ms_best'],4))"; }
run CSB200_RS_LB=1
timeout 300 python tools/next_rows_perf.py 2>&1 | grep -E "compress" | cut -c1-160
R='python tools/rmat_probe.py --scale 24 --iters 1 --no-gaxpy'
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_rs_\|k_bucket_hist --csv --log-file gpurun_out/r2y_8.csv $R > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2y_8.csv | grep "k_" | cut -c1-90
