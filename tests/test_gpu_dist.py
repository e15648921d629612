"""The sharded paths on the GPU: cs_multiply by column blocks of B and cs_gaxpy by
row blocks, against the oracle.  world = 1 runs in-process; world = 2 launches tools/dist_check.py
under torchrun over NCCL and needs two GPUs (skipped on a one-GPU box; the world-2 host logic is
covered on CPU by tests/test_dist_gloo.py)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

import csparse_cuda as cc
from csparse_cuda import synth, dist as csd
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("gen", [lambda: synth.st27(20), lambda: synth.rmat(11, 8)])
def test_sharded_multiply_world1(gen):
    """One rank owns every column: the gathered (Cp, Ci, Cx) is the whole product
    (csparse.py:1608-1642: pattern exact after the canonical sort, values 1e-12)."""
    import torch
    m, n, p, i, x = gen()
    A = orc.csc(m, n, p, i, x)
    R = orc.canonical(orc.cs_multiply(A, A))
    nr = int(R.p[R.n])
    dA = cc.from_arrays(m, n, p, i, x)
    bounds = csd.multiply_column_bounds(p, p, i, 1)
    assert list(bounds) == [0, n]
    for mode in ("root", "all"):
        dCl, got = csd.sharded_multiply(dA, dA, bounds, 0, gather=mode, device="cuda")
        Cp, Ci, Cx = (t.cpu().numpy() for t in got)
        Cz = orc.canonical(orc.csc(m, n, Cp, Ci, Cx))
        assert np.array_equal(Cp, R.p) and np.array_equal(Cz.i[:nr], R.i[:nr])
        assert np.all(np.abs(Cz.x[:nr] - R.x[:nr]) <= 1e-12 * np.abs(R.x[:nr]))
    # several blocks on one GPU, stitched by hand exactly as gather_columns places them
    bounds = csd.multiply_column_bounds(p, p, i, 3)
    parts = []
    for r in range(3):
        dCl, none = csd.sharded_multiply(dA, dA, bounds, r, gather=None, device="cuda")
        assert none is None
        parts.append(dCl.arrays())
    cps, off = [], 0
    for lp, _, _ in parts:
        cps.append(lp[:-1].astype(np.int64) + off)
        off += int(lp[-1])
    Cp = np.concatenate(cps + [np.array([off])]).astype(np.int32)
    Ci = np.concatenate([q[1][: q[0][-1]] for q in parts])
    Cx = np.concatenate([q[2][: q[0][-1]] for q in parts])
    Cz = orc.canonical(orc.csc(m, n, Cp, Ci, Cx))
    assert np.array_equal(Cp, R.p) and np.array_equal(Cz.i[:nr], R.i[:nr])
    assert np.all(np.abs(Cz.x[:nr] - R.x[:nr]) <= 1e-12 * np.abs(R.x[:nr]))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_paths_world2_nccl():
    """tools/dist_check.py under torchrun: halo (batched and fused) and all-gather cs_gaxpy, column-
    sharded cs_multiply with the final gather, every rank against the oracle."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tools", "dist_check.py"), "256"]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist_check: OK" in r.stdout
