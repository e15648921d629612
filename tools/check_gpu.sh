#!/bin/bash
# What every change of this round was checked with on the B200 box (gpurun -- 'bash tools/check_gpu.sh [1|2|N]'):
#   1 (default)  the whole GPU suite, the default bench line, the reference arm, smoke()
#   2            two GPUs: tests/test_gpu_dist.py, tools/dist_check.py at 1 M rows per rank, bench at N = 2 (fused and NCCL halo)
#   N >= 4       bench at N GPUs as the driver launches it
MODE=${1:-1}
mkdir -p gpurun_out
if [ "$MODE" = "1" ]; then
  python __graft_entry__.py --smoke 2>&1 | tail -1
  timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/check_pytest.log 2>&1; echo rc_pytest=$?; tail -3 gpurun_out/check_pytest.log
  timeout 900 python bench.py > gpurun_out/check_bench_default.json 2> gpurun_out/check_bench_default.err; echo rc_bench=$?
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/check_bench_reference.json 2>&1; echo rc_ref=$?
elif [ "$MODE" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_dist.py -x -q --timeout 800 -p no:cacheprovider 2>&1 | tail -3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 1024 > gpurun_out/check_dist_1024.log 2>&1; echo rc_dist=$?; grep "dist_check" gpurun_out/check_dist_1024.log
  for mode in fused nccl; do
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 300 --warmup 5 --halo $mode --no-extra > gpurun_out/check_bench_n2_$mode.json 2> gpurun_out/check_bench_n2_$mode.err; echo rc_bench_$mode=$?
  done
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $MODE --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $MODE --steps 100 --warmup 5 > gpurun_out/check_bench_n$MODE.json 2> gpurun_out/check_bench_n$MODE.err; echo rc_bench=$?
fi
