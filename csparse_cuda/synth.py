"""Deterministic synthetic matrices for the BASELINE.json configs.

Host-side numpy generators; every function returns ``(m, n, p, i, x)`` with
``p``/``i`` int32 and ``x`` float64, columns sorted and duplicate-free
("canonical CSC").  They only build inputs; no product arithmetic lives here.

* ``lap2d(k)``  -- 2-D 5-point Laplacian on a k x k grid (config C3: k=4096)
* ``st27(k)``   -- 3-D 27-point stencil on a k^3 grid (config C4: k=128)
* ``rmat(scale, ef)`` -- R-MAT power-law matrix, 2^scale rows (config C5: 24, 16)
"""
from __future__ import annotations

import numpy as np


def lap2d_cols(kx: int, ky: int, j0: int, j1: int):
    """Columns [j0, j1) of the 5-point Laplacian on a kx x ky grid (index = ix + kx*iy),
    with GLOBAL row indices.  The matrix is symmetric, so these are also rows [j0, j1)
    of its CSR view.  Returns (m, j1-j0, p, i, x)."""
    n = kx * ky
    j = np.arange(j0, j1, dtype=np.int64)
    ix = j % kx
    iy = j // kx
    rows = np.stack([j - kx, j - 1, j, j + 1, j + kx], axis=1)
    valid = np.stack([iy > 0, ix > 0, np.ones(len(j), bool), ix < kx - 1, iy < ky - 1], axis=1)
    vals = np.broadcast_to(np.array([-1.0, -1.0, 4.0, -1.0, -1.0]), (len(j), 5))
    p = np.zeros(len(j) + 1, dtype=np.int64)
    np.cumsum(valid.sum(axis=1), out=p[1:])
    i = rows[valid].astype(np.int32)
    x = np.ascontiguousarray(vals[valid], dtype=np.float64)
    return n, len(j), p.astype(np.int32), i, x


def lap2d(k: int):
    """A = I (x) L1 + L1 (x) I with L1 = tridiag(-1, 2, -1); index = ix + k*iy."""
    m, n, p, i, x = lap2d_cols(k, k, 0, k * k)
    assert p[-1] == 5 * n - 4 * k
    return m, n, p, i, x


def st27(k: int, seed: int = 0):
    """Pattern T (x) T (x) T with T = tridiag(1,1,1); values uniform(0.5, 1.5)."""
    n = k ** 3
    j = np.arange(n, dtype=np.int64)
    ix = j % k
    iy = (j // k) % k
    iz = j // (k * k)
    cols_rows = []
    cols_valid = []
    for dz in (-1, 0, 1):
        vz = (iz + dz >= 0) & (iz + dz < k)
        for dy in (-1, 0, 1):
            vy = (iy + dy >= 0) & (iy + dy < k)
            for dx in (-1, 0, 1):
                vx = (ix + dx >= 0) & (ix + dx < k)
                cols_rows.append(j + dx + k * dy + k * k * dz)
                cols_valid.append(vz & vy & vx)
    rows = np.stack(cols_rows, axis=1)
    valid = np.stack(cols_valid, axis=1)
    del cols_rows, cols_valid
    p = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(valid.sum(axis=1), out=p[1:])
    i = rows[valid].astype(np.int32)
    nnz = int(p[-1])
    assert nnz == (3 * k - 2) ** 3
    x = np.random.default_rng(seed).uniform(0.5, 1.5, nnz)
    return n, n, p.astype(np.int32), i, x


def rmat_edges(scale: int, ef: int = 16, a=0.57, b=0.19, c=0.19, seed: int = 1):
    """Raw R-MAT edge list (rows, cols, vals), duplicates still present."""
    rng = np.random.default_rng(seed)
    ne = ef << scale
    r = np.zeros(ne, dtype=np.int64)
    col = np.zeros(ne, dtype=np.int64)
    for lvl in range(scale):
        u = rng.random(ne)
        r |= (u >= a + b).astype(np.int64) << lvl
        col |= (((u >= a) & (u < a + b)) | (u >= a + b + c)).astype(np.int64) << lvl
    v = rng.uniform(-1.0, 1.0, ne)
    return r, col, v


def rmat(scale: int, ef: int = 16, a=0.57, b=0.19, c=0.19, seed: int = 1):
    """R-MAT as canonical CSC: duplicates summed (in edge order), columns sorted."""
    n = 1 << scale
    r, col, v = rmat_edges(scale, ef, a, b, c, seed)
    key = col * n + r
    order = np.argsort(key, kind="stable")
    key = key[order]
    v = v[order]
    first = np.ones(len(key), bool)
    first[1:] = key[1:] != key[:-1]
    starts = np.flatnonzero(first)
    x = np.add.reduceat(v, starts) if len(v) else v
    ukey = key[starts]
    i = (ukey % n).astype(np.int32)
    cj = ukey // n
    p = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(cj, minlength=n), out=p[1:])
    return n, n, p.astype(np.int32), i, np.ascontiguousarray(x, dtype=np.float64)


def rmat_torch(scale: int, ef: int = 16, a=0.57, b=0.19, c=0.19, seed: int = 1, device="cuda"):
    """The same R-MAT family generated on the GPU with torch (input generation only;
    scale 24 takes minutes in numpy).  Returns torch tensors (p int32, i int32, x float64)
    of the canonical CSC: duplicates summed, columns sorted.  Not bit-identical to rmat()."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    n = 1 << scale
    ne = ef << scale
    r = torch.zeros(ne, dtype=torch.int64, device=device)
    col = torch.zeros(ne, dtype=torch.int64, device=device)
    for lvl in range(scale):
        u = torch.rand(ne, generator=g, device=device, dtype=torch.float64)
        r |= (u >= a + b).to(torch.int64) << lvl
        col |= (((u >= a) & (u < a + b)) | (u >= a + b + c)).to(torch.int64) << lvl
        del u
    v = torch.rand(ne, generator=g, device=device, dtype=torch.float64) * 2.0 - 1.0
    key = col * n + r
    del r, col
    key, order = torch.sort(key)
    v = v[order]
    del order
    ukey, counts = torch.unique_consecutive(key, return_counts=True)
    del key
    ends = torch.cumsum(counts, 0)
    cs = torch.cumsum(v, 0)
    x = cs[ends - 1]
    x[1:] -= cs[ends[:-1] - 1]          # segment sums (inputs only: rounding of the generator is irrelevant)
    del cs, v, ends, counts
    i = (ukey % n).to(torch.int32)
    cj = ukey // n
    p = torch.zeros(n + 1, dtype=torch.int64, device=device)
    p[1:] = torch.cumsum(torch.bincount(cj, minlength=n), 0)
    return n, n, p.to(torch.int32), i, x.contiguous()


def vectors(m: int, n: int):
    """x = default_rng(0).standard_normal(n), y0 = default_rng(1).standard_normal(m)."""
    return (np.random.default_rng(0).standard_normal(n),
            np.random.default_rng(1).standard_normal(m))


def gaxpy_bytes(m: int, n: int, nnz: int) -> int:
    """Algorithmic bytes of cs_gaxpy: 12 nnz + 4(n+1) + 8n + 16m."""
    return 12 * nnz + 4 * (n + 1) + 8 * n + 16 * m


def transpose_bytes(m: int, n: int, nnz: int, values: bool = True) -> int:
    """24 nnz + 4(n+1) + 4(m+1) (pattern only: 8 nnz + ...)."""
    return (24 if values else 8) * nnz + 4 * (n + 1) + 4 * (m + 1)


def multiply_bytes(nnzA: int, nnzB: int, nnzC: int, nA: int, nB: int) -> int:
    """12(nnzA+nnzB+nnzC) + 4(n_A+1) + 8(n_B+1)."""
    return 12 * (nnzA + nnzB + nnzC) + 4 * (nA + 1) + 8 * (nB + 1)


def cumsum_bytes(n: int) -> int:
    return 12 * n + 4
