#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_mirror.py tests/test_gpu_parity.py -x -q --timeout 200 -p no:cacheprovider -k "transpose or mirror" 2>&1 | tail -1
Q='python tools/quick_perf.py --only transpose --lap 4096 --st 128 --rmat 0'
$Q 2>&1 | grep "transpose\[mirror\]" | cut -c1-120
