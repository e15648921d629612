#!/bin/bash
# two GPUs: the sharded paths against the oracle, then the bench at N = 2 (both halo modes)
timeout 900 python -m pytest tests/test_gpu_dist.py -x -q --timeout 800 -p no:cacheprovider 2>&1 | tail -6
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 1024 > gpurun_out/r2m_dist_check_1024.log 2>&1; echo rc_dc=$?; tail -4 gpurun_out/r2m_dist_check_1024.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 200 --warmup 5 > gpurun_out/r2m_bench_n2.json 2> gpurun_out/r2m_bench_n2.err; echo rc_bench=$?; tail -c 3000 gpurun_out/r2m_bench_n2.json; tail -5 gpurun_out/r2m_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 200 --warmup 5 --halo nccl --no-extra > gpurun_out/r2m_bench_n2_nccl.json 2> gpurun_out/r2m_bench_n2_nccl.err; echo rc_bench_nccl=$?; tail -c 1500 gpurun_out/r2m_bench_n2_nccl.json
