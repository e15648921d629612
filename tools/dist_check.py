#!/usr/bin/env python
"""torchrun check of the sharded cs_gaxpy on GPUs: every rank compares its slice of y with the CPU
oracle on the whole matrix over several steps with a changing x (halo exchange overlapped with the
interior rows).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csparse_cuda as cc
from csparse_cuda import synth, dist as csd
from oracle import oracle as orc

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cc.set_device(local)
cc.set_stream(torch.cuda.current_stream().cuda_stream)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ky = k * world
n = k * ky
bounds = csd.even_bounds(ky, world) * k
r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
m_, n_, p, i, x = synth.lap2d_cols(k, ky, r0, r1)
mf, nf, pf, i_f, xf = synth.lap2d_cols(k, ky, 0, n)       # the whole (symmetric) matrix for the oracle
A = orc.csc(n, n, pf, i_f, xf)
ok = True
for gated in (False,):
    blk = csd.RowBlock(r0, r1, p, i, x, int(i.min()), int(i.max()))
    sh = csd.ShardedGaxpy(blk, n, n, bounds, make_local=csd.cuda_make_local, local_spmv=csd.cuda_local_spmv,
                          device="cuda")
    rng = np.random.default_rng(5)
    y_ref = rng.standard_normal(n)
    y_own = torch.from_numpy(y_ref[r0:r1].copy()).cuda()
    xv = sh.own_view()
    for step in range(4):
        xg = rng.standard_normal(n)
        xv.copy_(torch.from_numpy(xg[r0:r1]).cuda())
        sh.step(xv, y_own)
        orc.cs_gaxpy(A, xg, y_ref)
        torch.cuda.synchronize()
        got = y_own.cpu().numpy()
        same = np.array_equal(got.view(np.int64), y_ref[r0:r1].view(np.int64))
        ok &= same
        if not same:
            print(f"rank {rank} gated={gated} step {step}: max err {np.abs(got - y_ref[r0:r1]).max():.3e}", flush=True)
    dist.barrier()
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_check:", "OK (bit-exact on every rank)" if int(t.item()) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
