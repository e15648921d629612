#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q --timeout 800 -p no:cacheprovider 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 1024 > gpurun_out/r2o_dist_check_1024.log 2>&1; echo rc_dc=$?; grep -E "dist_check|rank" gpurun_out/r2o_dist_check_1024.log | grep -v Traceback | tail -6
for mode in fused nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 300 --warmup 5 --halo $mode --no-extra > gpurun_out/r2o_bench_n2_$mode.json 2> gpurun_out/r2o_bench_n2_$mode.err; echo rc_bench_$mode=$?
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2o_bench_n2_$mode.json').read().strip().splitlines()[-1])
    print("$mode", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], d["gpu_launches"])
except Exception as e:
    print("$mode", "no line", e)
PY
done
