#!/bin/bash
# ncu evidence of round 2, final binary (each command first exits 0 without ncu; --clock-control none; one GPU).
# Summaries: python tools/ncu_summary.py {launches|raw} <csv> > profiles/r2_<name>.md ; python tools/prof_src.py <source csv>
mkdir -p gpurun_out
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto'
T='python tools/quick_perf.py --only transpose --lap 0 --rmat 0 --st 128 --once'
R='python tools/rmat_probe.py --scale 24 --iters 1 --no-gaxpy'
B2='python bench.py --steps 2 --warmup 3 --no-cpu'
$M > gpurun_out/p_plain_mm.log 2>&1 && ncu --set full --clock-control none -k regex:k_num_soa -s 1 -c 1 -o gpurun_out/r2f_num_soa -f $M > /dev/null 2>&1
$R > gpurun_out/p_plain_rmat.log 2>&1 && ncu --set full --clock-control none -k regex:k_rs_pass -s 3 -c 3 -o gpurun_out/r2f_rs_pass -f $R > /dev/null 2>&1
$T > gpurun_out/p_plain_tr.log 2>&1 && ncu --set full --clock-control none -k regex:k_bucket_sort -s 1 -c 1 -o gpurun_out/r2f_bucket_sort -f $T > /dev/null 2>&1
$R > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:k_ --log-file gpurun_out/r2f_rmat_tr_launches.csv $R > /dev/null 2>&1
$B2 > gpurun_out/p_plain_bench2.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2f_bench_launches.csv $B2 > /dev/null 2>&1
for n in num_soa rs_pass bucket_sort; do ncu -i gpurun_out/r2f_$n.ncu-rep --page raw --csv > gpurun_out/r2f_$n.raw.csv; ncu -i gpurun_out/r2f_$n.ncu-rep --page source --csv > gpurun_out/r2f_$n.source.csv; done
ls -la gpurun_out/r2f_*
