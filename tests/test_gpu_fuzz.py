"""Randomised GPU parity: many small matrices of awkward shapes (empty rows / columns, unsorted
columns, duplicates, explicit zeros, pattern-only, 1 x n, n x 1) through every entry point and both
algorithm choices where there are two, against the CPU oracle.  Deterministic seeds."""
import numpy as np
import pytest

import csparse_cuda as cc
from oracle import oracle as orc
from tests.test_gpu_parity import as_omat, assert_multiply_parity, assert_same_matrix, bits, to_cs

pytestmark = pytest.mark.gpu


def random_csc(rng, m, n, density, sort, dups, values=True):
    cols = []
    for _ in range(n):
        k = rng.binomial(m, density) if m else 0
        r = rng.integers(0, m, k) if dups else rng.choice(m, size=min(k, m), replace=False)
        if sort:
            r = np.sort(r) if dups else np.sort(r)
        cols.append(np.asarray(r, np.int32))
    p = np.zeros(n + 1, np.int32)
    p[1:] = np.cumsum([len(c) for c in cols])
    i = np.concatenate(cols).astype(np.int32) if p[-1] else np.zeros(0, np.int32)
    x = rng.standard_normal(len(i))
    x[rng.random(len(i)) < 0.1] = 0.0                      # explicit zeros
    return orc.csc(m, n, p, i if len(i) else np.zeros(1, np.int32)[:0], x if values else None)


def shapes(rng):
    m, n = int(rng.integers(1, 70)), int(rng.integers(1, 70))
    return m, n, float(rng.choice([0.0, 0.02, 0.1, 0.4, 0.9]))


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_transpose_gaxpy(seed):
    rng = np.random.default_rng(1000 + seed)
    m, n, d = shapes(rng)
    A = random_csc(rng, m, n, d, sort=bool(seed % 2), dups=bool(seed % 3 == 0), values=bool(seed % 5))
    for path in (None, "radix", "bucket"):
        cc.force_transpose_path(path)
        try:
            C = cc.cs_transpose(to_cs(A, lists=False), True)
        finally:
            cc.force_transpose_path(None)
        assert_same_matrix(C, orc.cs_transpose(A, True), f"transpose seed {seed} path {path}")
    if A.x is not None:
        xv, y0 = rng.standard_normal(n), rng.standard_normal(m)
        yr = y0.copy(); orc.cs_gaxpy(A, xv, yr)
        y = y0.copy(); assert cc.cs_gaxpy(to_cs(A, lists=False), xv, y)
        assert np.linalg.norm(y - yr) <= 1e-12 * max(np.linalg.norm(yr), 1e-300)


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_multiply_add(seed):
    rng = np.random.default_rng(2000 + seed)
    m, k, d = shapes(rng)
    n = int(rng.integers(1, 70))
    canon = bool(seed % 2)
    A = random_csc(rng, m, k, d, sort=canon, dups=not canon, values=bool(seed % 7))
    B = random_csc(rng, k, n, float(rng.choice([0.05, 0.3, 0.8])), sort=canon, dups=not canon, values=bool(seed % 4))
    R = orc.cs_multiply(A, B)
    assert_same_matrix(cc.cs_multiply(to_cs(A, lists=False), to_cs(B, lists=False)), R, f"ordered multiply seed {seed}")
    dA, dB = cc.upload(to_cs(A, lists=False)), cc.upload(to_cs(B, lists=False))
    for path in (None, "blocked_v1", "blocked_v2"):  # every version of the blocked numeric kernel
        cc.force_multiply_path(path)
        try:
            C = cc.cs_multiply(dA, dB).download(trim=True)
        finally:
            cc.force_multiply_path(None)
        assert_multiply_parity(C, R, f"device multiply seed {seed} path {path}")
        if R.x is not None and R.nnz:
            cz, rz = orc.canonical(as_omat(C)), orc.canonical(R)
            assert np.array_equal(bits(cz.x), bits(rz.x)), f"multiply values seed {seed} path {path}"
    # add: same shape operands
    A2 = random_csc(rng, m, k, d, sort=canon, dups=not canon, values=bool(seed % 3))
    al, be = float(rng.standard_normal()), float(rng.standard_normal())
    Ra = orc.cs_add(A, A2, al, be)
    for path in (None, "spgemm"):
        cc.force_add_path(path)
        try:
            Ca = cc.cs_add(to_cs(A, lists=False), to_cs(A2, lists=False), al, be)
        finally:
            cc.force_add_path(None)
        assert_same_matrix(Ca, Ra, f"add seed {seed} path {path}")


@pytest.mark.parametrize("seed", range(30))
def test_fuzz_assemble(seed):
    rng = np.random.default_rng(3000 + seed)
    m, n, d = shapes(rng)
    nz = int(rng.integers(0, 400))
    ti, tj = rng.integers(0, m, nz).astype(np.int32), rng.integers(0, n, nz).astype(np.int32)
    tx = rng.standard_normal(nz) if seed % 4 else None
    T = cc.cs()
    T.m, T.n, T.nz, T.nzmax, T.i, T.p, T.x = m, n, nz, max(nz, 1), ti, tj, tx
    R = orc.cs_compress(m, n, ti, tj, tx)
    C = cc.cs_compress(T)
    assert_same_matrix(C, R, f"compress seed {seed}")
    if tx is not None:
        R2 = R.copy(); orc.cs_dupl(R2)
        assert cc.cs_dupl(C) is True
        assert_same_matrix(C, R2, f"dupl seed {seed}")
        assert cc.cs_norm(to_cs(R)) == orc.cs_norm(R)
    A = random_csc(rng, m, n, d, sort=False, dups=True, values=bool(seed % 3))
    for mode, name in ((cc.KEEP_NONZERO, "nonzero"), (cc.KEEP_OFFDIAG, "dropdiag")):
        Rf = A.copy(); rn = orc.cs_fkeep(Rf, name)
        Af = to_cs(A, lists=False)
        assert cc.cs_fkeep(Af, mode, None) == rn
        assert_same_matrix(Af, Rf, f"fkeep {name} seed {seed}")
    pinv = np.empty(m, np.int32); pinv[rng.permutation(m)] = np.arange(m, dtype=np.int32)
    q = rng.permutation(n).astype(np.int32)
    assert_same_matrix(cc.cs_permute(to_cs(A, lists=False), pinv, q, True), orc.cs_permute(A, pinv, q, True), f"permute {seed}")
    if m == n or seed % 2:
        S = random_csc(rng, n, n, d, sort=False, dups=True, values=bool(seed % 2))
        pv = np.empty(n, np.int32); pv[rng.permutation(n)] = np.arange(n, dtype=np.int32)
        assert_same_matrix(cc.cs_symperm(to_cs(S, lists=False), pv, True), orc.cs_symperm(S, pv, True), f"symperm {seed}")


def test_concurrent_threads_on_different_handles():
    """SURVEY.md 8b threading: ctypes releases the GIL, so Python threads really overlap inside the
    library; calls on different handles must not disturb each other (thread-local stream / error
    state, per-call workspaces)."""
    import threading
    from csparse_cuda import synth
    errors = []

    def work(tid):
        try:
            rng = np.random.default_rng(50 + tid)
            m, n, p, i, x = synth.rmat(11 + tid % 3, 8) if tid % 2 else synth.lap2d(60 + 7 * tid)
            A = orc.csc(m, n, p, i, x)
            RT, RM = orc.cs_transpose(A, True), orc.cs_multiply(A, A)
            xv, y0 = rng.standard_normal(n), rng.standard_normal(m)
            yr = y0.copy(); orc.cs_gaxpy(A, xv, yr)
            for _ in range(8):
                dA = cc.from_arrays(m, n, p, i, x)
                tp, ti, tx = cc.cs_transpose(dA, True).arrays()
                assert np.array_equal(tp, RT.p) and np.array_equal(ti, RT.i[:A.nnz]) and np.array_equal(bits(tx), bits(RT.x[:A.nnz]))
                assert_same_matrix(cc.cs_multiply(to_cs(A, lists=False), to_cs(A, lists=False)), RM, f"thread {tid}")
                y = y0.copy(); cc.cs_gaxpy(dA, xv, y)
                assert np.linalg.norm(y - yr) <= 1e-12 * np.linalg.norm(yr)
                dA.free()
        except Exception as e:                      # noqa: BLE001
            errors.append((tid, repr(e)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


@pytest.mark.parametrize("path", [None, "blocked_v1", "blocked_v2"])
@pytest.mark.parametrize("canon", [True, False])
def test_blocked_multiply_long_columns(path, canon):
    """columns of A longer than a warp step (33..100 entries) that still give columns of C within the
    blocked kernel's 128 rows; chunks of B(:,j) longer than 32; duplicates inside A's columns"""
    rng = np.random.default_rng(77)
    A = random_csc(rng, 120, 90, 0.6, sort=canon, dups=not canon)
    B = random_csc(rng, 90, 60, 0.5, sort=canon, dups=not canon)
    R = orc.cs_multiply(A, B)
    dA, dB = cc.upload(to_cs(A, lists=False)), cc.upload(to_cs(B, lists=False))
    cc.force_multiply_path(path)
    try:
        C = cc.cs_multiply(dA, dB).download(trim=True)
    finally:
        cc.force_multiply_path(None)
    assert_multiply_parity(C, R, f"long columns path {path}")
    cz, rz = orc.canonical(as_omat(C)), orc.canonical(R)
    assert np.array_equal(bits(cz.x), bits(rz.x))
