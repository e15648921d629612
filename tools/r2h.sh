#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_templates.py -x -q --timeout 120 -p no:cacheprovider 2>&1 | tail -3
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto'
for v in 2 3 4 6; do CSB200_SOA_CTAS=$v timeout 200 $M 2>&1 | grep cs_multiply | sed "s/^/soa_ctas $v: /"; done > gpurun_out/r2h_variants.log
cat gpurun_out/r2h_variants.log
M1='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2h_mm_launches.csv $M1 > gpurun_out/r2h_ncu_mm.log 2>&1; echo rc_ncu=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_num_soa -s 1 -c 1 -o gpurun_out/r2h_num_soa -f $M1 > gpurun_out/r2h_ncu_full.log 2>&1; echo rc_full=$?
