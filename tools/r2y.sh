#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_next_rows.py tests/test_gpu_fuzz.py -x -q --timeout 200 -p no:cacheprovider -k "transpose or compress or symperm" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q --timeout 500 -p no:cacheprovider -k "c5" 2>&1 | tail -1
run() { env "$@" timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-gaxpy 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$*', round(d['ms_median'],4), round(d['ms_best'],4))"; }
run A=1
run A=2
timeout 300 python tools/next_rows_perf.py 2>&1 | grep compress | cut -c1-160
