#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mirror.py tests/test_gpu_fuzz.py tests/test_gpu_next_rows.py -x -q --timeout 120 -p no:cacheprovider -k "transpose or compress or fuzz or mirror or symperm or permute" 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q --timeout 280 -p no:cacheprovider -k "c3" 2>&1 | tail -3
timeout 300 python tools/quick_perf.py --only transpose --lap 4096 --st 0 --rmat 0 2>&1 | grep transpose | cut -c1-160
timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans split 2>&1 | tail -1 | cut -c1-160
CSB200_LONG_EF=0 timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans split 2>&1 | tail -1 | sed 's/^/no evict-first: /' | cut -c1-160
