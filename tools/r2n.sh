#!/bin/bash
# one GPU: the whole GPU suite, the default bench line, launch list of the bench, split SpMV timings
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/r2n_pytest.log 2>&1; echo rc_pytest=$?; tail -4 gpurun_out/r2n_pytest.log
timeout 900 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err; echo rc_bench=$?; tail -c 6000 gpurun_out/r2n_bench_default.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2n_bench_reference.json 2>&1; echo rc_ref=$?
timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-transpose --plans split,merge > gpurun_out/r2n_rmat.log 2>&1; tail -2 gpurun_out/r2n_rmat.log
timeout 300 python tools/rmat_probe.py --scale 24 --iters 3 --no-gaxpy > gpurun_out/r2n_rmat_tr.log 2>&1; tail -1 gpurun_out/r2n_rmat_tr.log
