# round 2, first call: evidence the round-1 verdict asked for, taken on the round-1 binary
set -x
R="python tools/rmat_probe.py --scale 24 --iters 1"
$R > gpurun_out/r2a_plain_rmat.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_spmv_merge" -s 2 -c 1 -o gpurun_out/r2a_merge -f $R > gpurun_out/r2a_ncu_merge.log 2>&1
echo rc_merge=$?
M="python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto"
$M > gpurun_out/r2a_plain_mm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_num_blocked2|k_sym_flat" -s 2 -c 2 -o gpurun_out/r2a_spgemm128 -f $M > gpurun_out/r2a_ncu_mm.log 2>&1
echo rc_mm=$?
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
