#!/bin/bash
timeout 300 python tools/quick_perf.py --only transpose --lap 4096 --st 128 --rmat 0 > gpurun_out/r2f_tr.log 2>&1; echo rc_tr=$?
cat gpurun_out/r2f_tr.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mirror.py tests/test_gpu_fuzz.py tests/test_gpu_next_rows.py -x -q --timeout 120 -p no:cacheprovider -k "transpose or compress or fuzz or mirror" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q --timeout 280 -p no:cacheprovider -k "c3" 2>&1 | tail -4
