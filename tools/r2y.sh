#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_templates.py tests/test_gpu_parity.py tests/test_gpu_next_rows.py -x -q --timeout 200 -p no:cacheprovider -k "multiply or template or tpl or amd or add" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q --timeout 500 -p no:cacheprovider -k "c4" 2>&1 | tail -1
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto'
$M 2>&1 | grep multiply | cut -c1-130
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2y_mm.csv $M --once > /dev/null 2>&1
python tools/ncu_summary.py launches gpurun_out/r2y_mm.csv | grep "k_" | cut -c1-90 | head -6
