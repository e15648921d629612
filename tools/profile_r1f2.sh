set -x
M="python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 96 --once --mul-paths auto"
$M > gpurun_out/r1f_plain_sym.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_sym_flat" -s 0 -c 3 -o gpurun_out/r1f_sym -f $M > gpurun_out/r1f_ncu_sym.log 2>&1
echo rc_sym=$?
