"""world_size-2 gloo tests of the sharding layer (csparse_cuda/dist.py): the same
partition / halo / all-gather code the NCCL path runs, with the local SpMV injected
as a CPU function so that no GPU is needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from csparse_cuda import dist as csd
from csparse_cuda import synth
from oracle import oracle as orc


def test_balanced_bounds_and_plans():
    rp = np.array([0, 10, 10, 10, 20, 40, 40, 41])
    b = csd.balanced_bounds(rp, 2)
    assert b[0] == 0 and b[-1] == 7 and np.all(np.diff(b) >= 0)
    assert abs(int(rp[b[1]]) - 20) <= 10
    assert csd.balanced_bounds(np.zeros(5, np.int64), 3).tolist()[0] == 0
    assert csd.even_bounds(10, 4).tolist() == [0, 2, 5, 7, 10]
    xb = np.array([0, 100, 200, 300])
    pl = csd.plan_exchange(1, 3, xb, [(0, 110), (90, 215), (180, 299)], 300)
    assert pl.mode == "halo" and (pl.win_lo, pl.win_hi) == (90, 216) and pl.lo_need == [0, 10, 20]
    assert pl.hi_need == [11, 16, 0]
    pl = csd.plan_exchange(0, 3, xb, [(0, 250), (90, 215), (180, 299)], 300)   # reaches past the neighbour
    assert pl.mode == "gather" and (pl.win_lo, pl.win_hi) == (0, 300)


def test_multiply_column_bounds():
    m, n, p, i, x = synth.st27(5)
    b = csd.multiply_column_bounds(p, p, i, 4)
    assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) > 0)
    lens = np.diff(p)[i]
    per_col = np.add.reduceat(lens, p[:-1])
    loads = [per_col[b[g]:b[g + 1]].sum() for g in range(4)]
    assert max(loads) <= 1.2 * (sum(loads) / 4)


def _cpu_make_local(rowptr, col_local, val, ncols_local):
    # CSC of the transposed block == CSR of the block
    return orc.csc(ncols_local, len(rowptr) - 1, rowptr, col_local, val)


def _cpu_local_spmv(AT, x_window, y_own):
    # y += AT' * x : row-by-row over the CSR view
    A = orc.cs_transpose(AT, True)               # (rows x window) as CSC
    y = y_own.numpy()
    orc.cs_gaxpy(A, np.ascontiguousarray(x_window.numpy()), y)


def _worker(rank, world, port, kind, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if kind == "lap2d":
            m, n, p, i, x = synth.lap2d(24)
        else:
            m, n, p, i, x = synth.rmat(9, 6)
        A = orc.csc(m, n, p, i, x)
        AT = orc.cs_transpose(A, True)           # CSR view of A
        bounds = csd.balanced_bounds(AT.p.astype(np.int64), world)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
        blk = csd.csr_row_block(AT.p, AT.i, AT.x, r0, r1)
        sh = csd.ShardedGaxpy(blk, m, n, bounds, make_local=_cpu_make_local, local_spmv=_cpu_local_spmv,
                              device="cpu", force_gather=(kind == "rmat"))
        xv, y0 = synth.vectors(m, n)
        x_own = torch.from_numpy(xv[r0:r1].copy())
        y_own = torch.from_numpy(y0[r0:r1].copy())
        sh.step(x_own, y_own)
        sh.step(x_own, y_own)                    # second step: buffers are reusable
        yref = y0.copy()
        orc.cs_gaxpy(A, xv, yref)
        orc.cs_gaxpy(A, xv, yref)
        err = float(np.abs(y_own.numpy() - yref[r0:r1]).max())
        q.put((rank, sh.plan.mode, err, sh.exchanged_bytes, r1 - r0))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("kind,mode", [("lap2d", "halo"), ("rmat", "gather")])
def test_sharded_gaxpy_world2(kind, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, got_mode, err, nbytes, rows in res:
        assert got_mode == mode, res
        assert err <= 1e-12, res
        if mode == "halo":
            assert nbytes == 8 * 24             # one grid line of the 24 x 24 Laplacian per neighbour


def _col_block(B, j0, j1):
    b, e = int(B.p[j0]), int(B.p[j1])
    return orc.csc(B.m, j1 - j0, (B.p[j0:j1 + 1] - b).astype(np.int32), B.i[b:e].copy(), B.x[b:e].copy())


def _worker_mul(rank, world, port, kind, q):
    """Column-sharded cs_multiply: every rank forms C(:, J_rank) = A * B(:, J_rank) (here with the
    oracle standing in for the local GPU product) and the final gather assembles C -- on rank 0
    ("root") and on every rank ("all").  The gathered (Cp, Ci, Cx) must equal the whole product."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m, n, p, i, x = synth.st27(6) if kind == "st27" else synth.rmat(8, 5)
        A = orc.csc(m, n, p, i, x)
        bounds = csd.multiply_column_bounds(p, p, i, world)
        if kind == "rmat":
            bounds = np.array([0, 0, n][: world + 1] if world == 2 else bounds)   # an EMPTY block on rank 0
        j0, j1 = int(bounds[rank]), int(bounds[rank + 1])
        Cl = orc.cs_multiply(A, _col_block(A, j0, j1))
        nl = int(Cl.p[Cl.n])
        cp, ci, cx = torch.from_numpy(Cl.p.copy()), torch.from_numpy(Cl.i[:nl].copy()), torch.from_numpy(Cl.x[:nl].copy())
        Cref = orc.cs_multiply(A, A)
        nr = int(Cref.p[Cref.n])
        ok = True
        for mode in ("root", "all"):
            got = csd.gather_columns(cp, ci, cx, bounds, rank, world, mode, None, "cpu")
            if mode == "root" and rank != 0:
                ok &= got is None
                continue
            Cp, Ci, Cx = (t.numpy() for t in got)
            ok &= np.array_equal(Cp, Cref.p) and np.array_equal(Ci, Cref.i[:nr])
            ok &= np.array_equal(Cx.view(np.int64), Cref.x[:nr].view(np.int64))
        # pattern-only pieces travel without a value array
        got = csd.gather_columns(cp, ci, None, bounds, rank, world, "all", None, "cpu")
        ok &= got[2] is None and np.array_equal(got[1].numpy(), Cref.i[:nr])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["st27", "rmat"])
def test_sharded_multiply_gather_world2(kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_mul, args=(r, 2, port, kind, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res
