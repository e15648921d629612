#!/bin/bash
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto'
for v in 0 1 2; do CSB200_TPL_VARIANT=$v timeout 200 $M 2>&1 | grep cs_multiply | sed "s/^/variant $v: /"; done > gpurun_out/r2e_variants.log
cat gpurun_out/r2e_variants.log
timeout 300 python -m pytest tests/test_gpu_templates.py -x -q --timeout 60 -p no:cacheprovider 2>&1 | tail -3
M1='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2e_mm_launches.csv $M1 > gpurun_out/r2e_ncu_mm.log 2>&1; echo rc_ncu=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_num_tpl -s 1 -c 1 -o gpurun_out/r2e_num_tpl -f $M1 > gpurun_out/r2e_ncu_full.log 2>&1; echo rc_full=$?
