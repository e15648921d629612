#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_templates.py tests/test_gpu_parity.py -x -q --timeout 120 -p no:cacheprovider -k "templates or multiply or gaxpy" 2>&1 | tail -5
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto'
for v in 3 4; do CSB200_SOA_CTAS=$v timeout 200 $M 2>&1 | grep cs_multiply | sed "s/^/soa_ctas $v: /"; done > gpurun_out/r2i_variants.log
cat gpurun_out/r2i_variants.log
M1='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2i_mm_launches.csv $M1 > gpurun_out/r2i_ncu_mm.log 2>&1; echo rc_ncu=$?
timeout 400 python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans auto,merge > gpurun_out/r2i_rmat.log 2>&1; echo rc_rmat=$?; cat gpurun_out/r2i_rmat.log | tail -4
