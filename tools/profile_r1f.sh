set -x
T="python tools/quick_perf.py --only transpose --st 0 --rmat 0 --lap 4096 --once"
$T > gpurun_out/r1f_plain_tr.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_mirror" -s 1 -c 1 -o gpurun_out/r1f_mirror -f $T > gpurun_out/r1f_ncu_tr.log 2>&1
echo rc_tr=$?
M="python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 96 --once --mul-paths blocked_v2"
$M > gpurun_out/r1f_plain_mm.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_num_blocked2|k_sym_flat" -s 2 -c 2 -o gpurun_out/r1f_spgemm -f $M > gpurun_out/r1f_ncu_mm.log 2>&1
echo rc_mm=$?
