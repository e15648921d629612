#!/bin/bash
M='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --mul-paths auto'
for v in 4 5 6; do CSB200_SOA_CTAS=$v timeout 200 $M 2>&1 | grep cs_multiply | sed "s/^/soa_ctas $v: /"; done > gpurun_out/r2p_variants.log
cat gpurun_out/r2p_variants.log
M1='python tools/quick_perf.py --only multiply --lap 0 --rmat 0 --st 128 --once --mul-paths auto'
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2p_mm_launches.csv $M1 > gpurun_out/r2p_ncu_mm.log 2>&1; echo rc_ncu=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_num_soa -s 1 -c 1 -o gpurun_out/r2p_num_soa -f $M1 > gpurun_out/r2p_ncu_full.log 2>&1; echo rc_full=$?
# transpose bucket path: warp-per-bucket kernel with 16 entries per lane on the 27-point stencil
T='python tools/quick_perf.py --only transpose --lap 4096 --st 128 --rmat 0'
timeout 300 $T 2>&1 | grep "bucket\]" | sed "s/^/default: /" > gpurun_out/r2p_tr.log
NVCC_EXTRA="-DWB_EPT_DEF=16" timeout 600 python -m csparse_cuda.build --force > /dev/null 2>&1; echo rc_build=$?
CSB200_TR_WARP_AVG=40 timeout 300 $T 2>&1 | grep "bucket\]" | sed "s/^/ept16 warp_avg40: /" >> gpurun_out/r2p_tr.log
timeout 300 $T 2>&1 | grep "bucket\]" | sed "s/^/ept16 warp_avg8: /" >> gpurun_out/r2p_tr.log
CSB200_TR_WARP_AVG=40 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q --timeout 120 -p no:cacheprovider -k "transpose" 2>&1 | tail -2 >> gpurun_out/r2p_tr.log
cat gpurun_out/r2p_tr.log
