#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_templates.py tests/test_gpu_parity.py -x -q --timeout 200 -p no:cacheprovider -k "multiply or template or tpl" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q --timeout 500 -p no:cacheprovider -k "c4" 2>&1 | tail -1
for st in 0 1 2 4; do echo stride=$st; CSB200_SOA_STRIDE=$st python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto 2>&1 | grep multiply | cut -c1-130; done
