#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_templates.py tests/test_gpu_parity.py -x -q --timeout 120 -p no:cacheprovider -k "templates or multiply or gaxpy" 2>&1 | tail -4
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto'
timeout 200 $M 2>&1 | grep cs_multiply > gpurun_out/r2l_mm.log; cat gpurun_out/r2l_mm.log
M1='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_mm_launches.csv $M1 > gpurun_out/r2l_ncu_mm.log 2>&1; echo rc_ncu=$?
timeout 400 python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans split,merge > gpurun_out/r2l_rmat.log 2>&1; echo rc_rmat=$?; tail -3 gpurun_out/r2l_rmat.log
R='python tools/rmat_probe.py --scale 24 --iters 1 --no-transpose --plans split'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2l_rmat_launches.csv $R > gpurun_out/r2l_ncu_rmat.log 2>&1; echo rc_ncu_rmat=$?
