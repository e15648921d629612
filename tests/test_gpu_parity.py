"""GPU parity tests: csparse_cuda (through the C ABI of libcsparse_b200.so) against
the CPU oracle on the same inputs, and against the golden vectors made from the
unmodified reference.

Contract (BASELINE.json north_star):
  cs_transpose, cs_cumsum : bit-exact p, i, x
  cs_multiply             : bit-exact pattern after canonical per-column sort,
                            values within 1e-12 relative (we additionally check
                            the stronger property that p/i/x match the reference's
                            discovery order bit for bit on canonical inputs)
  cs_gaxpy                : 1e-12 normwise relative error (the row-stream kernel
                            is additionally bit-exact)
"""
import numpy as np
import pytest

import csparse_cuda as cc
from csparse_cuda import synth
from oracle import oracle as orc
from tests.golden_util import ALL, FIXTURES, KNOWN, Golden, assert_values_close, edge_cases, normwise

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def to_cs(M, lists=True):
    """oracle OMat -> csparse_cuda.cs (list-backed like the reference, or numpy-backed)."""
    A = cc.cs()
    A.m, A.n, A.nz, A.nzmax = M.m, M.n, M.nz, M.nzmax
    if lists:
        A.p, A.i = M.p.tolist(), M.i.tolist()
        A.x = None if M.x is None else M.x.tolist()
    else:
        A.p, A.i, A.x = M.p, M.i, M.x
    return A


def as_omat(Ccs):
    return orc.OMat(Ccs.m, Ccs.n, np.asarray(Ccs.p, np.int32), np.asarray(Ccs.i, np.int32),
                    None if Ccs.x is None else np.asarray(Ccs.x, np.float64), Ccs.nzmax, Ccs.nz)


def bits(a):
    return np.asarray(a, np.float64).view(np.int64)


def assert_same_matrix(C, R, what=""):
    """bit-exact p / i / x and identical shape conventions"""
    assert (C.m, C.n, C.nz, C.nzmax) == (R.m, R.n, R.nz, R.nzmax), what
    assert len(C.p) == len(R.p) and len(C.i) == len(R.i), what
    assert np.array_equal(np.asarray(C.p), np.asarray(R.p)), what + " p"
    assert np.array_equal(np.asarray(C.i), np.asarray(R.i)), what + " i"
    if R.x is None:
        assert C.x is None, what
    else:
        assert C.x is not None and len(C.x) == len(R.x), what
        assert np.array_equal(bits(C.x), bits(R.x)), what + " x"


def assert_multiply_parity(C, R, what=""):
    """the north_star contract for cs_multiply"""
    assert (C.m, C.n, C.nz, C.nzmax) == (R.m, R.n, R.nz, R.nzmax), what
    assert np.array_equal(np.asarray(C.p), np.asarray(R.p)), what + " p"
    cc_, rr = orc.canonical(as_omat(C)), orc.canonical(R)
    assert np.array_equal(cc_.i, rr.i), what + " pattern after canonical sort"
    if R.x is None:
        assert C.x is None
    else:
        assert_values_close(cc_.x, rr.x, RTOL, what=what + " values")


# ---- cs_transpose ------------------------------------------------------------------

@pytest.mark.parametrize("name", ALL)
def test_transpose_fixtures(name):
    g = Golden(name)
    A = g.A()
    for lists in (True, False):
        C = cc.cs_transpose(to_cs(A, lists), True)
        assert_same_matrix(C, orc.cs_transpose(A, True), name)
        g.check("AT", as_omat(C))
    Cp = cc.cs_transpose(to_cs(A), False)
    assert Cp.x is None
    g.check("ATpattern", as_omat(Cp))
    ATT = cc.cs_transpose(cc.cs_transpose(to_cs(A), True), True)
    g.check("ATT", as_omat(ATT))


def test_transpose_known_norms():
    """1-norms the reference's CSparseTest1 asserts for A' (csparse_test.py:269-426)."""
    for name in FIXTURES:
        _, (m, n, nnz, _, _), (nrmT, dT), _ = KNOWN[name]
        AT = cc.cs_transpose(to_cs(Golden(name).A()), True)
        assert (AT.m, AT.n, AT.p[AT.n]) == (n, m, nnz)
        assert abs(orc.cs_norm(as_omat(AT)) - nrmT) <= dT


@pytest.mark.parametrize("gen", [lambda: synth.lap2d(300), lambda: synth.st27(20),
                                 lambda: synth.rmat(13, 16), lambda: synth.rmat(16, 16)])
def test_transpose_synthetic(gen):
    m, n, p, i, x = gen()
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_transpose(A, True)
    dA = cc.from_arrays(m, n, p, i, x)
    dC = cc.cs_transpose(dA, True)
    cp, ci, cx = dC.arrays()
    assert np.array_equal(cp, R.p) and np.array_equal(ci, R.i[:A.nnz])
    assert np.array_equal(bits(cx), bits(R.x[:A.nnz]))
    # transposing twice returns the (already sorted) input bit for bit
    p2, i2, x2 = cc.cs_transpose(dC, True).arrays()
    assert np.array_equal(p2, p) and np.array_equal(i2, i) and np.array_equal(bits(x2), bits(x))


def test_transpose_slab_schedule():
    """The opt-in schedules of the bucket sort -- partition and sort interleaved per L2-sized slab
    (csb200_transpose_force_path(3), needs >= 4 M entries to cut anything) and fused into one persistent
    launch (force_path 4): same bits as the oracle."""
    m, n, p, i, x = synth.lap2d(1024)
    x = np.random.default_rng(8).standard_normal(len(i))
    R = orc.cs_transpose(orc.csc(m, n, p, i, x), True)
    dA = cc.from_arrays(m, n, p, i, x)
    for path in ("bucket_slab", "bucket_fused", "bucket"):
        cc.force_transpose_path(path)
        try:
            cp, ci, cx = cc.cs_transpose(dA, True).arrays()
            assert cc.last_transpose_path() == "bucket"
        finally:
            cc.force_transpose_path(None)
        assert np.array_equal(cp, R.p) and np.array_equal(ci, R.i[:len(i)]), path
        assert np.array_equal(bits(cx), bits(R.x[:len(i)])), path


def test_transpose_unsorted_duplicates_long_rows():
    """Unsorted columns, duplicate (i,j) entries and rows far longer than a warp/CTA tile."""
    rng = np.random.default_rng(7)
    m, n = 40, 3000
    cols = []
    for j in range(n):
        k = int(rng.integers(0, 30))
        r = rng.integers(0, m, k)          # with repetition: duplicates inside the column, unsorted
        cols.append(r)
    p = np.zeros(n + 1, np.int64)
    p[1:] = np.cumsum([len(c) for c in cols])
    i = np.concatenate(cols).astype(np.int32)
    x = rng.standard_normal(len(i))
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_transpose(A, True)            # rows of A hold ~2000 entries each
    C = cc.cs_transpose(to_cs(A, lists=False), True)
    assert_same_matrix(C, R, "dups/long")
    # one very long row next to short ones
    m2, n2 = 5, 20000
    i2 = np.where(rng.random(n2) < 0.9, 0, rng.integers(1, m2, n2)).astype(np.int32)
    p2 = np.arange(n2 + 1)
    x2 = rng.standard_normal(n2)
    A2 = orc.csc(m2, n2, p2, i2, x2)
    assert_same_matrix(cc.cs_transpose(to_cs(A2, lists=False), True), orc.cs_transpose(A2, True), "long row")


@pytest.fixture
def radix_path():
    cc.force_transpose_path("radix")
    yield
    cc.force_transpose_path(None)


@pytest.mark.parametrize("name", ALL)
def test_transpose_radix_fixtures(name, radix_path):
    """the stable radix sort (power-law path) on every fixture: same bits as the reference"""
    test_transpose_fixtures(name)


def test_transpose_radix_synthetic(radix_path):
    for gen in (lambda: synth.lap2d(300), lambda: synth.st27(20), lambda: synth.rmat(16, 16)):
        test_transpose_synthetic(gen)
    test_transpose_unsorted_duplicates_long_rows()
    # pattern only, > 1 pass, ragged tail tile
    m, n, p, i, x = synth.rmat(17, 8)
    A = orc.csc(m, n, p, i, None)
    C = cc.cs_transpose(to_cs(A, lists=False), False)
    assert_same_matrix(C, orc.cs_transpose(A, False), "radix pattern")


def test_transpose_power_law_takes_radix_path_and_matches():
    """R-MAT rows overflow the row buckets: the automatic choice must still be bit-exact"""
    m, n, p, i, x = synth.rmat(18, 16)
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_transpose(A, True)
    cp, ci, cx = cc.cs_transpose(cc.from_arrays(m, n, p, i, x), True).arrays()
    assert np.array_equal(cp, R.p) and np.array_equal(ci, R.i[:A.nnz])
    assert np.array_equal(bits(cx), bits(R.x[:A.nnz]))


def test_transpose_deterministic():
    m, n, p, i, x = synth.rmat(14, 16)
    dA = cc.from_arrays(m, n, p, i, x)
    a = cc.cs_transpose(dA, True).arrays()
    b = cc.cs_transpose(dA, True).arrays()
    assert all(np.array_equal(u.view(np.uint8), v.view(np.uint8)) for u, v in zip(a, b))


# ---- cs_cumsum ----------------------------------------------------------------------

@pytest.mark.parametrize("n", [0, 1, 5, 2047, 2048, 2049, 4096, 100_003, 3_000_001])
def test_cumsum_sizes(n):
    rng = np.random.default_rng(n)
    c0 = rng.integers(0, 50, n).astype(np.int32)
    c_ref, p_ref = c0.copy(), np.full(n + 3, -7, np.int32)
    tot_ref = orc.cs_cumsum(p_ref, c_ref, n)
    c, p = c0.tolist() + [99], [-7] * (n + 3)
    tot = cc.cs_cumsum(p, c, n)
    assert tot == tot_ref == int(c0.sum())
    assert p == p_ref.tolist()            # tail beyond n untouched
    assert c[:n] == c_ref.tolist() and c[n] == 99
    # numpy-backed in-place form
    cn, pn = c0.copy(), np.zeros(n + 1, np.int32)
    assert cc.cs_cumsum(pn, cn, n) == tot_ref
    assert np.array_equal(pn, p_ref[: n + 1]) and np.array_equal(cn, c_ref)


def test_cumsum_golden_and_sentinels():
    for name in ALL:
        g = Golden(name)
        c = g.z["cumsum_in"].tolist()
        p = [7] * (len(c) + 1)
        assert cc.cs_cumsum(p, c, len(c)) == g.meta["cumsum_total"]
        assert p == g.z["cumsum_p"].tolist() and c == g.z["cumsum_c"].tolist()
    e = edge_cases()
    assert cc.cs_cumsum(None, [1], 1) == e["cumsum_none"] == -1
    assert cc.cs_cumsum([0, 0], None, 1) == -1
    p, c = [9] * 6, [3, 0, 2, 5, 11]
    assert cc.cs_cumsum(p, c, 4) == e["cumsum_t1_ret"]
    assert p == e["cumsum_t1_p"] and c == e["cumsum_t1_c"]
    p0 = [5]
    assert cc.cs_cumsum(p0, [], 0) == 0 and p0 == [0]


# ---- cs_gaxpy ------------------------------------------------------------------------

@pytest.mark.parametrize("name", ALL)
def test_gaxpy_fixtures(name):
    g = Golden(name)
    A = g.A()
    for M, key in ((A, "gaxpy_y"), (orc.cs_transpose(A, True), "gaxpy_yT")):
        x, y0 = synth.vectors(M.m, M.n)
        y = y0.tolist()
        assert cc.cs_gaxpy(to_cs(M), x.tolist(), y) is True
        ref = g.z[key]
        assert normwise(y, ref) <= RTOL
        # both plans, on a device-resident handle
        for plan in ("stream", "merge", "stream_ld", "split"):
            dA = cc.from_arrays(M.m, M.n, M.p, M.i, M.x)
            dA.force_gaxpy_plan(plan)
            yy = y0.copy()
            assert cc.cs_gaxpy(dA, x, yy) is True
            assert dA.gaxpy_plan() == plan
            assert normwise(yy, ref) <= RTOL, (name, plan)
            # sequential in-row order, no FMA: bit-exact -- except where a block of rows
            # overflows the shared-memory stage and falls back to warp-per-row (mbeacxc's
            # 250..484-entry rows, the power-law rows of rmat_9)
            if plan not in ("merge", "split") and name not in ("mbeacxc", "rmat_9"):
                assert np.array_equal(bits(yy), bits(ref)), (name, "stream plan not bit-exact")


@pytest.mark.parametrize("gen,plan", [(lambda: synth.lap2d(300), "stream"), (lambda: synth.st27(20), "stream"),
                                      (lambda: synth.rmat(16, 16), None), (lambda: synth.rmat(13, 4), None)])
def test_gaxpy_synthetic(gen, plan):
    m, n, p, i, x = gen()
    A = orc.csc(m, n, p, i, x)
    xv, y0 = synth.vectors(m, n)
    yref = y0.copy()
    orc.cs_gaxpy(A, xv, yref)
    dA = cc.from_arrays(m, n, p, i, x)
    if plan is not None:
        assert dA.gaxpy_plan() == plan       # the automatic choice
    for force in ("stream", "merge", "stream_ld", "split"):
        dA.force_gaxpy_plan(force)
        y = y0.copy()
        assert cc.cs_gaxpy(dA, xv, y)
        assert normwise(y, yref) <= RTOL, force
        if force == "stream" and dA.gaxpy_plan() == "stream":
            pass
    # repeated application accumulates: y0 + 2 A x
    y = y0.copy()
    cc.cs_gaxpy(dA, xv, y)
    cc.cs_gaxpy(dA, xv, y)
    y2 = yref.copy()
    orc.cs_gaxpy(A, xv, y2)
    assert normwise(y, y2) <= RTOL


def test_gaxpy_sentinels_and_shapes():
    A = to_cs(Golden("t1").A())
    T = cc.cs(); T.nz = 3
    assert cc.cs_gaxpy(T, [1.0], [1.0]) is False
    assert cc.cs_gaxpy(None, [1.0], [1.0]) is False
    assert cc.cs_gaxpy(A, None, [0.0] * 4) is False
    assert cc.cs_gaxpy(A, [1.0] * 4, None) is False
    P = to_cs(Golden("t1").A()); P.x = None
    with pytest.raises(TypeError):
        cc.cs_gaxpy(P, [1.0] * 4, [0.0] * 4)
    # the SURVEY golden vector; ints promoted; longer-than-needed x / y keep their tails
    y = [1, 2, 3, 4, 77]
    assert cc.cs_gaxpy(A, [1, 2, 3, 4, 55], y) is True
    assert y == [15.100000000000001, 14.499999999999998, 15.4, 12.3, 77]
    e = edge_cases()
    d = e["transpose_dups_in"]
    D = cc.cs(); D.m, D.n, D.nz, D.nzmax, D.p, D.i, D.x = d["m"], d["n"], -1, d["nzmax"], d["p"], d["i"], d["x"]
    yd = [0.5, -1.5, 2.5]
    assert cc.cs_gaxpy(D, [2.0, -3.0], yd)
    assert yd == e["gaxpy_dups_y"]
    # empty matrix: True, y untouched
    E = cc.cs(); E.m, E.n, E.nz, E.nzmax, E.p, E.i, E.x = 3, 2, -1, 1, [0, 0, 0], [0], [0.0]
    ye = [1.0, 2.0, 3.0]
    assert cc.cs_gaxpy(E, [1.0, 1.0], ye) is True and ye == [1.0, 2.0, 3.0]


# ---- cs_multiply ----------------------------------------------------------------------

@pytest.mark.parametrize("name", ALL)
def test_multiply_fixtures(name):
    g = Golden(name)
    A = g.A()
    AT = orc.cs_transpose(A, True)
    R = orc.cs_multiply(A, AT)
    C = cc.cs_multiply(to_cs(A), to_cs(AT))
    assert_multiply_parity(C, R, name)
    g.check("C", as_omat(C), order="pattern")
    assert_same_matrix(C, R, name + " discovery order")      # stronger than the contract
    # pattern-only product A'A as cs_amd forms it (csparse.py:250-254)
    ATp = orc.cs_transpose(A, False)
    Cp = cc.cs_multiply(to_cs(ATp), to_cs(A))
    assert Cp.x is None
    g.check("CpatternATA", as_omat(Cp), order="pattern")
    # the reference test's known answers: nnz and 1-norm of D = C + norm(C) I
    if name in KNOWN:
        _, (m, _, _, _, _), _, (dnnz, dnrm, dd) = KNOWN[name]
        Co = as_omat(C)
        eye = orc.csc(m, m, np.arange(m + 1), np.arange(m), np.ones(m))
        D = orc.cs_add(Co, eye, 1.0, orc.cs_norm(Co))
        assert D.nnz == dnnz and abs(orc.cs_norm(D) - dnrm) <= dd


@pytest.mark.parametrize("name", ["bcsstk01", "bcsstk16"])
def test_multiply_symmetrised(name):
    """BASELINE config 2: S*S' on the make_sym'ed matrix (csparse_test.py:115-121)."""
    g = Golden(name)
    S = orc.make_sym(g.A())
    ST = orc.cs_transpose(S, True)
    dS, dST = cc.upload(to_cs(S, False)), cc.upload(to_cs(ST, False))
    dC = cc.cs_multiply(dS, dST)
    C = dC.download(trim=True)
    assert_multiply_parity(C, orc.cs_multiply(S, ST), name)
    g.check("SST", as_omat(C), order="pattern")
    g.check("ST", as_omat(cc.cs_transpose(dS, True).download()))


@pytest.mark.parametrize("gen", [lambda: synth.lap2d(200), lambda: synth.st27(16), lambda: synth.rmat(11, 8)])
def test_multiply_synthetic_AA(gen):
    m, n, p, i, x = gen()
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_multiply(A, A)
    dA = cc.from_arrays(m, n, p, i, x)
    C = cc.cs_multiply(dA, dA).download(trim=True)          # device operands: blocked numeric kernel allowed
    assert_multiply_parity(C, R)
    cz, rz = orc.canonical(as_omat(C)), orc.canonical(R)
    assert np.array_equal(bits(cz.x), bits(rz.x)), "values are summed in the reference's sequence"
    cc.force_multiply_path("ordered")
    try:
        C = cc.cs_multiply(dA, dA).download(trim=True)
    finally:
        cc.force_multiply_path(None)
    assert_same_matrix(C, R, "discovery order")              # stronger than the contract
    A_host = to_cs(A, lists=False)
    assert_same_matrix(cc.cs_multiply(A_host, A_host), R, "host operands: reference order")


def test_multiply_large_columns_dense_path():
    """Columns of C beyond the shared-memory classes (> 1024 rows) use the dense workspaces."""
    rng = np.random.default_rng(3)
    m, k, n = 6000, 400, 37
    import scipy.sparse as sp
    Asp = sp.random(m, k, density=0.05, format="csc", random_state=11, data_rvs=rng.standard_normal)
    Bsp = sp.random(k, n, density=0.3, format="csc", random_state=12, data_rvs=rng.standard_normal)
    Asp.sort_indices(); Bsp.sort_indices()
    A = orc.csc(m, k, Asp.indptr, Asp.indices, Asp.data)
    B = orc.csc(k, n, Bsp.indptr, Bsp.indices, Bsp.data)
    R = orc.cs_multiply(A, B)
    assert np.diff(R.p).max() > 1024
    C = cc.cs_multiply(to_cs(A, False), to_cs(B, False))
    assert_multiply_parity(C, R, "dense path")
    assert_same_matrix(C, R, "dense path discovery order")
    # same with unsorted / duplicated A columns (non-canonical): tolerance contract only
    perm = rng.permutation(A.nnz)
    cols = np.repeat(np.arange(k), np.diff(A.p))
    order = np.lexsort((perm, cols))
    A2 = orc.csc(m, k, A.p, A.i[order], A.x[order])
    R2 = orc.cs_multiply(A2, B)
    C2 = cc.cs_multiply(to_cs(A2, False), to_cs(B, False))
    assert_multiply_parity(C2, R2, "dense path, unsorted A")


def test_multiply_edge_cases():
    e = edge_cases()

    def mk(d):
        A = cc.cs()
        A.m, A.n, A.nz, A.nzmax, A.p, A.i, A.x = d["m"], d["n"], d["nz"], d["nzmax"], d["p"], d["i"], d["x"]
        return A

    def same(C, d):
        assert (C.m, C.n, C.nz, C.nzmax) == (d["m"], d["n"], d["nz"], d["nzmax"])
        assert C.p == d["p"] and C.i == d["i"]
        if d["x"] is None:
            assert C.x is None
        else:
            assert np.array_equal(bits(C.x), bits(d["x"]))

    E32 = cc.cs(); E32.m, E32.n, E32.nz, E32.nzmax, E32.p, E32.i, E32.x = 3, 2, -1, 1, [0, 0, 0], [0], [0.0]
    E23 = cc.cs(); E23.m, E23.n, E23.nz, E23.nzmax, E23.p, E23.i, E23.x = 2, 3, -1, 1, [0, 0, 0, 0], [0], [0.0]
    same(cc.cs_transpose(E32, True), e["transpose_empty_3x2"])
    same(cc.cs_multiply(E23, E32), e["multiply_empty_2x3_3x2"])
    R = cc.cs(); R.m, R.n, R.nz, R.nzmax, R.p, R.i, R.x = 1, 2, -1, 2, [0, 1, 2], [0, 0], [1.0, -1.0]
    Cc = cc.cs(); Cc.m, Cc.n, Cc.nz, Cc.nzmax, Cc.p, Cc.i, Cc.x = 2, 1, -1, 2, [0, 2], [0, 1], [1.0, 1.0]
    same(cc.cs_multiply(R, Cc), e["multiply_cancel"])            # structural zero kept
    T = cc.cs(); T.nz = 1; T.m = T.n = 2; T.p, T.i, T.x = [0], [0], [1.0]
    assert cc.cs_transpose(T, True) is None and cc.cs_multiply(T, T) is None
    assert cc.cs_transpose(None, True) is None and cc.cs_multiply(None, R) is None
    assert cc.cs_multiply(R, R) is None                           # A.n != B.m
    Dm = mk(e["transpose_dups_in"])
    DT = cc.cs_transpose(Dm, True)
    same(DT, e["transpose_dups"])
    same(cc.cs_multiply(Dm, DT), e["multiply_dups"])
    Sp = cc.cs(); Sp.m, Sp.n, Sp.nz, Sp.nzmax = 2, 2, -1, 4
    Sp.p, Sp.i, Sp.x = [0, 2, 4], [0, 1, 0, 1], [-0.0, float("nan"), float("inf"), 5e-324]
    tsp = cc.cs_transpose(Sp, True)
    assert [int(v) for v in bits(tsp.x)] == e["transpose_special_x_bits"] and tsp.i == e["transpose_special_i"]
    Pn = cc.cs(); Pn.m, Pn.n, Pn.nz, Pn.nzmax, Pn.p, Pn.i, Pn.x = 3, 2, -1, 3, [0, 2, 3], [0, 2, 1], None
    same(cc.cs_transpose(Pn, True), e["transpose_pattern_only"])
    same(cc.cs_multiply(Pn, cc.cs_transpose(Pn, False)), e["multiply_pattern_only"])
    Lg = cc.cs(); Lg.m, Lg.n, Lg.nz, Lg.nzmax = 2, 2, -1, 4
    Lg.p, Lg.i, Lg.x = [0, 1, 2], [0, 1, 1, 0], [1.5, 2.5, 99.0, 98.0]
    same(cc.cs_transpose(Lg, True), e["transpose_tail_ignored"])
    # malformed contents are rejected, never written through
    Bad = cc.cs(); Bad.m, Bad.n, Bad.nz, Bad.nzmax, Bad.p, Bad.i, Bad.x = 2, 2, -1, 2, [0, 1, 2], [0, 5], [1.0, 1.0]
    with pytest.raises(ValueError):
        cc.cs_transpose(Bad, True)


def test_monkeypatched_reference_callers():
    """cs_compress-style caller: the stable scatter driven by cs_cumsum's cursor array
    (csparse.py:663-671) works with our cs_cumsum patched in."""
    g = Golden("west0067")
    Ti, Tj, Tx = g.z["T_i"], g.z["T_j"], g.z["T_x"]
    n = g.meta["T"]["n"]
    w = np.bincount(Tj, minlength=n).astype(np.int32).tolist()
    Cp = [0] * (n + 1)
    cc.cs_cumsum(Cp, w, n)
    Ci, Cx = [0] * len(Ti), [0.0] * len(Ti)
    for k in range(len(Ti)):
        q = w[Tj[k]]; w[Tj[k]] += 1
        Ci[q], Cx[q] = int(Ti[k]), float(Tx[k])
    A = g.A()
    assert Cp == A.p.tolist() and Ci == A.i.tolist() and Cx == A.x.tolist()


@pytest.mark.parametrize("name", ["bcsstk01", "west0067", "ash219"])
def test_numpy_backed_operands_give_numpy_backed_results(name):
    """A cs whose p / i / x are numpy arrays comes back numpy-backed (no per-element conversion);
    the same call on a list-backed cs gives lists with the same contents and shape conventions."""
    M = Golden(name).A()
    An, Al = to_cs(M, lists=False), to_cs(M, lists=True)
    Tn, Tl = cc.cs_transpose(An, True), cc.cs_transpose(Al, True)
    assert isinstance(Tn.i, np.ndarray) and isinstance(Tn.p, np.ndarray) and isinstance(Tn.x, np.ndarray)
    assert isinstance(Tl.i, list) and isinstance(Tl.p, list) and isinstance(Tl.x, list)
    assert Tn.i.dtype == np.int32 and Tn.x.dtype == np.float64
    assert_same_matrix(Tn, Tl, name + " transpose")
    Cn, Cl = cc.cs_multiply(An, Tn), cc.cs_multiply(Al, Tl)
    assert isinstance(Cn.i, np.ndarray) and isinstance(Cl.i, list)
    assert_same_matrix(Cn, Cl, name + " multiply")
    if M.m == M.n:
        Sn, Sl = cc.cs_add(An, Tn, 1.0, 2.0), cc.cs_add(Al, Tl, 1.0, 2.0)
        assert isinstance(Sn.i, np.ndarray) and isinstance(Sl.i, list)
        assert_same_matrix(Sn, Sl, name + " add")
