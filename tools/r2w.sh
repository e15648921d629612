#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_next_rows.py tests/test_gpu_mirror.py -x -q --timeout 200 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q --timeout 500 -p no:cacheprovider -k "c5" 2>&1 | tail -3
timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-gaxpy 2>&1 | tail -1 | cut -c1-160
CSB200_RS_GATHER=1 timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-gaxpy 2>&1 | tail -1 | sed 's/^/gather version: /' | cut -c1-160
timeout 300 python tools/next_rows_perf.py 2>&1 | tail -8 | cut -c1-160
R='python tools/rmat_probe.py --scale 24 --iters 1 --no-gaxpy'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2w_rmat_tr_launches.csv $R > gpurun_out/r2w_ncu.log 2>&1; echo rc_ncu=$?
