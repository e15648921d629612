#!/usr/bin/env python
"""torchrun check of the sharded paths on GPUs (NCCL), every rank against the CPU oracle on the
WHOLE matrix:

  * cs_gaxpy, banded matrix, halo mode: batched send/recv overlapped with the interior rows, and the
    fused form (halo lines pulled over NVLink inside the one SpMV launch, csb200_gaxpy_halo_dev) --
    bit-exact slices of y over several steps with a changing x;
  * cs_gaxpy, R-MAT, all-gather mode: 1e-12 normwise (csparse.py:1199-1213);
  * cs_multiply, column blocks of B with the final gather to rank 0 and to every rank: p exact,
    pattern exact after the canonical per-column sort, values 1e-12 (csparse.py:1608-1642).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_check.py [k]

tests/test_gpu_dist.py launches it when the box has at least two GPUs.
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
k = int(sys.argv[1]) if len(sys.argv) > 1 else 512
ok = True


def say(msg):
    print(f"[rank {rank}] {msg}", flush=True)


# ---- cs_gaxpy, halo mode ---------------------------------------------------------------------------
ky = k * world
n = k * ky
bounds = csd.even_bounds(ky, world) * k
r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
m_, n_, p, i, x = synth.lap2d_cols(k, ky, r0, r1)
mf, nf, pf, i_f, xf = synth.lap2d_cols(k, ky, 0, n)       # the whole (symmetric) matrix for the oracle
A = orc.csc(n, n, pf, i_f, xf)
for fused in (False, True):
    blk = csd.RowBlock(r0, r1, p, i, x, int(i.min()), int(i.max()))
    sh = csd.ShardedGaxpy(blk, n, n, bounds, make_local=csd.cuda_make_local, local_spmv=csd.cuda_local_spmv,
                          device="cuda", fused=fused)
    assert sh.plan.mode == "halo" and sh.fused == fused
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
            say(f"halo fused={fused} step {step}: max err {np.abs(got - y_ref[r0:r1]).max():.3e}")
    if fused:
        # the same step on pinned HOST slices (csb200_gaxpy_halo: chunked duplex copies around the halo pull)
        for step in range(2):
            xg = rng.standard_normal(n)
            hx = torch.from_numpy(xg[r0:r1].copy()).pin_memory()
            hy = torch.from_numpy(y_ref[r0:r1].copy()).pin_memory()
            sh.step_host(hx, hy)
            orc.cs_gaxpy(A, xg, y_ref)
            same = np.array_equal(hy.numpy().view(np.int64), y_ref[r0:r1].view(np.int64))
            ok &= same
            if not same:
                say(f"halo host step {step}: max err {np.abs(hy.numpy() - y_ref[r0:r1]).max():.3e}")
        # y resident on the device, the host gets a copy of the result
        y_dev = torch.from_numpy(y_ref[r0:r1].copy()).cuda()
        hy = torch.zeros(r1 - r0, dtype=torch.float64).pin_memory()
        for step in range(2):
            xg = rng.standard_normal(n)
            hx = torch.from_numpy(xg[r0:r1].copy()).pin_memory()
            sh.step_host(hx, hy, y_dev)
            orc.cs_gaxpy(A, xg, y_ref)
            same = np.array_equal(hy.numpy().view(np.int64), y_ref[r0:r1].view(np.int64)) and \
                np.array_equal(y_dev.cpu().numpy().view(np.int64), y_ref[r0:r1].view(np.int64))
            ok &= same
            if not same:
                say(f"halo host step {step} (y resident): max err {np.abs(hy.numpy() - y_ref[r0:r1]).max():.3e}")
        ok &= not sh.halo.timed_out()
    dist.barrier()
    if fused:
        sh.halo.free()
del A

# ---- cs_gaxpy, all-gather mode (R-MAT) -------------------------------------------------------------------
m, n, p, i, x = synth.rmat(13, 8)
A = orc.csc(m, n, p, i, x)
T = orc.cs_transpose(A, True)                               # CSR view
nnzT = int(T.p[T.n])
row_bounds = csd.balanced_bounds(T.p.astype(np.int64), world)
r0, r1 = int(row_bounds[rank]), int(row_bounds[rank + 1])
blk = csd.csr_row_block(T.p, T.i[:nnzT], T.x[:nnzT], r0, r1)
sh = csd.ShardedGaxpy(blk, m, n, row_bounds, x_bounds=csd.even_bounds(n, world), make_local=csd.cuda_make_local,
                      local_spmv=csd.cuda_local_spmv, device="cuda", force_gather=True)   # two ranks are always neighbours
assert sh.plan.mode == "gather"
c0, c1 = sh.c0, sh.c1
rng = np.random.default_rng(9)
y_ref = rng.standard_normal(m)
y_own = torch.from_numpy(y_ref[r0:r1].copy()).cuda()
for step in range(3):
    xg = rng.standard_normal(n)
    sh.step(torch.from_numpy(xg[c0:c1].copy()).cuda(), y_own)
    orc.cs_gaxpy(A, xg, y_ref)
torch.cuda.synchronize()
err = np.linalg.norm(y_own.cpu().numpy() - y_ref[r0:r1]) / max(np.linalg.norm(y_ref[r0:r1]), 1e-300)
if not err <= 1e-12:
    ok = False
    say(f"gather-mode cs_gaxpy: normwise error {err:.3e}")
dist.barrier()

# ---- cs_multiply, column blocks of B, final gather -----------------------------------------------------------
for name, (m, n, p, i, x) in (("st27 20^3", synth.st27(20)), ("rmat 11", synth.rmat(11, 8))):
    A = orc.csc(m, n, p, i, x)
    R = orc.canonical(orc.cs_multiply(A, A))
    nr = int(R.p[R.n])
    dA = cc.from_arrays(m, n, p, i, x)
    cbounds = csd.multiply_column_bounds(p, p, i, world)
    for mode in ("root", "all"):
        dCl, got = csd.sharded_multiply(dA, dA, cbounds, rank, gather=mode, device="cuda")
        if mode == "root" and rank != 0:
            ok &= got is None
            continue
        Cp, Ci, Cx = (t.cpu().numpy() for t in got)
        Cz = orc.canonical(orc.csc(m, n, Cp, Ci, Cx))
        good = np.array_equal(Cp, R.p) and np.array_equal(Cz.i[:nr], R.i[:nr])
        good = good and bool(np.all(np.abs(Cz.x[:nr] - R.x[:nr]) <= 1e-12 * np.abs(R.x[:nr])))
        if not good:
            ok = False
            say(f"sharded cs_multiply {name} gather={mode}: differs from the oracle")
    # C left column-distributed: every rank's block against the oracle's columns
    dCl, _ = csd.sharded_multiply(dA, dA, cbounds, rank, gather=None, device="cuda")
    j0, j1 = int(cbounds[rank]), int(cbounds[rank + 1])
    lp, li, lx = dCl.arrays()
    Lz = orc.canonical(orc.csc(m, j1 - j0, lp, li, lx))
    b, e = int(R.p[j0]), int(R.p[j1])
    good = np.array_equal(lp, R.p[j0:j1 + 1] - b) and np.array_equal(Lz.i[: e - b], R.i[b:e])
    if not good:
        ok = False
        say(f"sharded cs_multiply {name}: the local block differs from the oracle")
    dist.barrier()

t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("dist_check:", "OK (every rank matches the oracle)" if int(t.item()) else "FAILED", flush=True)
dist.destroy_process_group()
sys.exit(0 if int(t.item()) else 1)
