set -x
B="python bench.py --steps 20 --warmup 3 --no-cpu"
$B > gpurun_out/r1b_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1b_bench_launches.csv $B > gpurun_out/r1b_ncu_bench.log 2>&1
echo rc1=$?
M="python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 96 --once"
$M > gpurun_out/r1b_plain_mm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_num_warp|k_sym_flat" -s 2 -c 2 -o gpurun_out/r1b_spgemm -f $M > gpurun_out/r1b_ncu_mm.log 2>&1
echo rc2=$?
R="python tools/rmat_probe.py --scale 22 --iters 1"
$R > gpurun_out/r1b_plain_rmat.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_rs_pass|k_spmv_merge" -s 6 -c 4 -o gpurun_out/r1b_rmat -f $R > gpurun_out/r1b_ncu_rmat.log 2>&1
echo rc3=$?
T="python tools/quick_perf.py --only transpose --st 0 --rmat 0 --lap 4096 --once"
$T > gpurun_out/r1b_plain_tr.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_partition|k_bucket_sort_warp" -s 2 -c 2 -o gpurun_out/r1b_transpose -f $T > gpurun_out/r1b_ncu_tr.log 2>&1
echo rc4=$?
