#!/bin/bash
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto'
timeout 200 $M > gpurun_out/r2d_plain_mm.log 2>&1; echo rc_plain=$?
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_mm_launches.csv $M > gpurun_out/r2d_ncu_mm.log 2>&1; echo rc_ncu=$?
timeout 900 python -m pytest tests -x -m gpu --timeout 300 -v -p no:cacheprovider > gpurun_out/r2d_pytest.log 2>&1; echo rc_all=$?
tail -8 gpurun_out/r2d_pytest.log
