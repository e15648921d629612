#!/bin/bash
# ncu evidence of round 2 (each command first exits 0 without ncu; --clock-control none; one GPU).
# Summaries: python tools/ncu_summary.py {launches|raw} <csv> > profiles/r2_<name>.md
set -x
mkdir -p gpurun_out
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto'
R='python tools/rmat_probe.py --scale 24 --iters 1 --no-transpose --plans auto'
B='python bench.py --steps 3 --warmup 3 --no-extra --no-cpu'
B2='python bench.py --steps 2 --warmup 3 --no-cpu'
$M > gpurun_out/p_plain_mm.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_multiply_launches.csv $M > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_num_soa -s 1 -c 1 -o gpurun_out/r2_num_soa -f $M > /dev/null 2>&1
$R > gpurun_out/p_plain_rmat.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_spmv_long -s 2 -c 1 -o gpurun_out/r2_spmv_long -f $R > /dev/null 2>&1
$B > gpurun_out/p_plain_bench.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_spmv_tma -s 4 -c 1 -o gpurun_out/r2_spmv_tma -f $B > /dev/null 2>&1
$B2 > gpurun_out/p_plain_bench2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_bench_launches.csv $B2 > /dev/null 2>&1
for n in num_soa spmv_long spmv_tma; do ncu -i gpurun_out/r2_$n.ncu-rep --page raw --csv > gpurun_out/r2_$n.raw.csv; done
