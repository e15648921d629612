#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_templates.py tests/test_gpu_parity.py tests/test_gpu_next_rows.py -x -q --timeout 200 -p no:cacheprovider 2>&1 | tail -3
timeout 200 python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto 2>&1 | grep cs_multiply | cut -c1-200
timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans auto 2>&1 | tail -1 | cut -c1-200
# evidence: traffic of the headline kernel, full capture of the long-row kernel, launch list of the default bench
B='python bench.py --steps 3 --warmup 3 --no-extra --no-cpu'
timeout 300 $B > gpurun_out/r2t_bench_small.json 2>&1; echo rc_small=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_tma -s 4 -c 1 -o gpurun_out/r2t_spmv_tma -f $B > gpurun_out/r2t_ncu_tma.log 2>&1; echo rc_tma=$?
R='python tools/rmat_probe.py --scale 24 --iters 1 --no-transpose --plans auto'
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_spmv_long -s 2 -c 1 -o gpurun_out/r2t_spmv_long -f $R > gpurun_out/r2t_ncu_long.log 2>&1; echo rc_long=$?
B2='python bench.py --steps 2 --warmup 3 --no-cpu'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2t_bench_launches.csv $B2 > gpurun_out/r2t_ncu_bench.log 2>&1; echo rc_launches=$?
