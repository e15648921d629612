"""Full-size BASELINE configs on the GPU: compared with the C oracle where it finishes
in seconds, and through size-independent properties otherwise."""
import numpy as np
import pytest

import csparse_cuda as cc
from csparse_cuda import synth
from oracle import oracle as orc
from tests.golden_util import normwise

pytestmark = pytest.mark.gpu


def bits(a):
    return np.asarray(a, np.float64).view(np.int64)


def test_c3_lap2d_4096_transpose_and_gaxpy():
    """Config C3: n = 16 777 216, nnz = 83 869 696."""
    m, n, p, i, x = synth.lap2d(4096)
    assert len(i) == 83869696
    # unsymmetric values on the symmetric pattern, so that A' != A: the one-pass mirror path and the
    # two-hop bucket sort against the oracle
    xr = np.random.default_rng(3).standard_normal(len(i))
    R = orc.cs_transpose(orc.csc(m, n, p, i, xr), True)
    dR = cc.from_arrays(m, n, p, i, xr)
    for path, took in ((None, "mirror"), ("bucket", "bucket")):
        cc.force_transpose_path(path)
        try:
            tp, ti, tx = cc.cs_transpose(dR, True).arrays()
        finally:
            cc.force_transpose_path(None)
        assert cc.last_transpose_path() == took
        assert np.array_equal(tp, R.p) and np.array_equal(ti, R.i) and np.array_equal(bits(tx), bits(R.x)), path
    del R, xr, tx
    dR.free()
    A = orc.csc(m, n, p, i, x)
    dA = cc.from_arrays(m, n, p, i, x)
    tp, ti, tx = cc.cs_transpose(dA, True).arrays()
    # symmetric matrix: A' == A bit for bit; and transposing twice is the identity
    assert np.array_equal(tp, p) and np.array_equal(ti, i) and np.array_equal(bits(tx), bits(x))
    xv, y0 = synth.vectors(m, n)
    yref = y0.copy()
    orc.cs_gaxpy(A, xv, yref)
    for plan in ("stream", "merge", "split"):
        dA.force_gaxpy_plan(plan)
        y = y0.copy()
        assert cc.cs_gaxpy(dA, xv, y)
        assert normwise(y, yref) <= 1e-12
        if plan == "stream":
            assert np.array_equal(bits(y), bits(yref))        # every block fits a stage: bit-exact
    # linearity: A(2x) accumulates exactly twice A x on top of y0 = 0
    z1, z2 = np.zeros(m), np.zeros(m)
    cc.cs_gaxpy(dA, xv, z1)
    cc.cs_gaxpy(dA, 2.0 * xv, z2)
    assert np.array_equal(bits(2.0 * z1), bits(z2))


def test_c4_st27_multiply():
    """Config C4: full oracle comparison at 64^3 (every kernel version) and at the full 128^3."""
    m, n, p, i, x = synth.st27(64)
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_multiply(A, A)
    dA = cc.from_arrays(m, n, p, i, x)
    assert R.nnz == (5 * 64 - 6) ** 3
    # default device path (blocked numeric kernel): the contract -- pattern after canonical sort, values
    Rz = orc.canonical(R)
    for path in (None, "blocked_v1"):
        cc.force_multiply_path(path)
        try:
            dC = cc.cs_multiply(dA, dA)
        finally:
            cc.force_multiply_path(None)
        cp, ci, cx = dC.arrays()
        Cz = orc.canonical(orc.csc(m, n, cp, ci, cx))
        assert np.array_equal(cp, R.p) and np.array_equal(Cz.i, Rz.i), path
        assert np.array_equal(bits(Cz.x), bits(Rz.x)), path                # same summation sequence: bit-equal
        del dC, Cz
    del Rz
    cc.force_multiply_path("ordered")
    try:
        cp, ci, cx = cc.cs_multiply(dA, dA).arrays()
    finally:
        cc.force_multiply_path(None)
    assert np.array_equal(cp, R.p) and np.array_equal(ci, R.i[:R.nnz])      # discovery order
    assert np.array_equal(bits(cx), bits(R.x[:R.nnz]))
    del R, cp, ci, cx
    dA.free()
    # ---- the BASELINE configuration itself: 128^3, nnz(C) = 254 840 104, against the oracle ----------
    m, n, p, i, x = synth.st27(128)
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_multiply(A, A)                        # csparse.py:1608-1642, ~10 s in C
    nr = int(R.p[R.n])
    assert nr == 634 ** 3 == 254840104
    dA = cc.from_arrays(m, n, p, i, x)
    dC = cc.cs_multiply(dA, dA)
    assert dC.nnz == nr
    assert cc.last_multiply_flops() == 1142 ** 3 == 1489355288
    cp, ci, cx = dC.arrays()
    assert np.array_equal(cp, R.p)
    if cc.last_multiply_templated() == n:
        # pattern-class templates: the reference's discovery order and summation sequence, bit for bit
        assert np.array_equal(ci, R.i[:nr]) and np.array_equal(bits(cx), bits(R.x[:nr]))
    else:
        Cz, Rz = orc.canonical(orc.csc(m, n, cp, ci, cx)), orc.canonical(R)
        assert np.array_equal(Cz.i[:nr], Rz.i[:nr])                      # pattern after the canonical sort
        assert np.all(np.abs(Cz.x[:nr] - Rz.x[:nr]) <= 1e-12 * np.abs(Rz.x[:nr]))
        del Cz, Rz
    del cp, ci, cx
    # the general kernels (no templates) at the same size: the contract
    cc.force_multiply_path("no_templates")
    try:
        dG = cc.cs_multiply(dA, dA)
    finally:
        cc.force_multiply_path(None)
    assert cc.last_multiply_templated() == 0
    gp, gi, gx = dG.arrays()
    Gz, Rz = orc.canonical(orc.csc(m, n, gp, gi, gx)), orc.canonical(R)
    assert np.array_equal(gp, R.p) and np.array_equal(Gz.i[:nr], Rz.i[:nr])
    assert np.all(np.abs(Gz.x[:nr] - Rz.x[:nr]) <= 1e-12 * np.abs(Rz.x[:nr]))
    del Gz, Rz, gp, gi, gx, dG, R, A
    # C v == A (A v) within rounding (a checksum of checksums), v > 0 so no cancellation
    v = np.random.default_rng(5).uniform(0.5, 1.5, n)
    t = np.zeros(m); cc.cs_gaxpy(dA, v, t)
    u = np.zeros(m); cc.cs_gaxpy(dA, t, u)
    w = np.zeros(m); cc.cs_gaxpy(dC, v, w)
    assert normwise(w, u) <= 1e-12
    # transpose of C has the same column counts (A*A is structurally symmetric)
    tp = cc.cs_transpose(dC, False).arrays()[0]
    assert np.array_equal(tp, dC.arrays()[0])


def test_c5_rmat20_transpose_and_gaxpy():
    """Config C5 family at scale 20 (16 M nnz, rows up to ~40 k entries): merge-path plan."""
    m, n, p, i, x = synth.rmat(20, 16)
    A = orc.csc(m, n, p, i, x)
    dA = cc.from_arrays(m, n, p, i, x)
    tp, ti, tx = cc.cs_transpose(dA, True).arrays()
    R = orc.cs_transpose(A, True)
    assert np.array_equal(tp, R.p) and np.array_equal(ti, R.i) and np.array_equal(bits(tx), bits(R.x))
    xv, y0 = synth.vectors(m, n)
    yref = y0.copy()
    orc.cs_gaxpy(A, xv, yref)
    y = y0.copy()
    assert cc.cs_gaxpy(dA, xv, y)
    assert dA.gaxpy_plan() == "split"
    assert normwise(y, yref) <= 1e-12
    untouched = np.diff(R.p) == 0                 # empty rows keep y bit for bit
    assert np.array_equal(bits(y[untouched]), bits(y0[untouched]))
    dA.force_gaxpy_plan("merge")
    y = y0.copy()
    assert cc.cs_gaxpy(dA, xv, y) and dA.gaxpy_plan() == "merge"
    assert normwise(y, yref) <= 1e-12


def test_c5_rmat24_transpose_and_gaxpy():
    """Config C5 at full size: R-MAT scale 24 (n = 16 777 216, ~263 M nnz), generated on the GPU,
    downloaded once, cs_transpose bit-exact and cs_gaxpy 1e-12 against the oracle
    (csparse.py:2292-2315, :1199-1213)."""
    import torch
    m, n, tp, ti, tx = synth.rmat_torch(24, 16)
    nnz = int(ti.numel())
    assert nnz > 260_000_000
    dA = cc.from_device(m, n, tp.data_ptr(), ti.data_ptr(), tx.data_ptr())
    torch.cuda.synchronize()
    p, i, x = tp.cpu().numpy(), ti.cpu().numpy(), tx.cpu().numpy()
    del tp, ti, tx
    torch.cuda.empty_cache()
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_transpose(A, True)
    dT = cc.cs_transpose(dA, True)
    assert cc.last_transpose_path() == "radix"
    cp, ci, cx = dT.arrays()
    assert np.array_equal(cp, R.p)
    assert np.array_equal(ci, R.i[:nnz])
    assert np.array_equal(bits(cx), bits(R.x[:nnz]))
    del cp, ci, cx
    dT.free()
    xv, y0 = synth.vectors(m, n)
    yref = y0.copy()
    orc.cs_gaxpy(A, xv, yref)
    y = y0.copy()
    assert cc.cs_gaxpy(dA, xv, y)
    assert dA.gaxpy_plan() == "split"
    assert normwise(y, yref) <= 1e-12
    untouched = np.diff(R.p) == 0                 # empty rows keep y bit for bit
    assert np.array_equal(bits(y[untouched]), bits(y0[untouched]))
