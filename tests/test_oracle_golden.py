"""Pins the C oracle (oracle/csparse_oracle.c) bit-for-bit against vectors made by
the unmodified Python reference (oracle/make_golden.py) and against the known
answers of the reference's own CSparseTest1 (csparse_test.py:269-426)."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.golden_util import ALL, FIXTURES, KNOWN, KNOWN_SYM, Golden, edge_cases


@pytest.mark.parametrize("name", ALL)
def test_compress_transpose_multiply_flow(name):
    g = Golden(name)
    if "T" in g.meta:  # cs_compress (csparse.py:647) from the triplets as loaded
        t = g.meta["T"]
        A = orc.cs_compress(t["m"], t["n"], g.z["T_i"], g.z["T_j"], g.z["T_x"])
        g.check("A", A)
    A = g.A()
    AT = orc.cs_transpose(A, True)
    g.check("AT", AT)
    assert AT.nzmax == g.meta["AT"]["nzmax"] == max(A.nnz, 1)
    if g.has("AT"):
        ref = g.mat("AT")
        assert np.array_equal(AT.p, ref.p) and np.array_equal(AT.i[:A.nnz], ref.i)
        assert np.array_equal(AT.x[:A.nnz].view(np.int64), ref.x.view(np.int64))
    ATp = orc.cs_transpose(A, False)
    assert ATp.x is None
    g.check("ATpattern", ATp)
    g.check("ATT", orc.cs_transpose(AT, True))
    C = orc.cs_multiply(A, AT)
    g.check("C", C)                       # discovery order, bit-exact values
    assert C.nzmax == g.meta["C"]["nzmax"] == g.meta["C"]["nnz"]
    g.check("CpatternATA", orc.cs_multiply(ATp, A))
    eye = orc.csc(A.m, A.m, np.arange(A.m + 1), np.arange(A.m), np.ones(A.m))
    D = orc.cs_add(C, eye, 1.0, orc.cs_norm(C))
    g.check("D", D)
    assert orc.cs_norm(C) == g.meta["C"]["norm1"]
    if "AA" in g.meta:
        g.check("AA", orc.cs_multiply(A, A))


@pytest.mark.parametrize("name", ALL)
def test_gaxpy_cumsum_dupl(name):
    from csparse_cuda import synth
    g = Golden(name)
    A = g.A()
    x, y = synth.vectors(A.m, A.n)
    assert orc.cs_gaxpy(A, x, y) is True
    assert np.array_equal(y.view(np.int64), g.z["gaxpy_y"].view(np.int64))
    AT = orc.cs_transpose(A, True)
    xt, yt = synth.vectors(A.n, A.m)
    assert orc.cs_gaxpy(AT, xt, yt)
    assert np.array_equal(yt.view(np.int64), g.z["gaxpy_yT"].view(np.int64))
    c = g.z["cumsum_in"].copy()
    p = np.full(A.m + 1, 7, np.int32)
    assert orc.cs_cumsum(p, c, A.m) == g.meta["cumsum_total"]
    assert np.array_equal(p, g.z["cumsum_p"]) and np.array_equal(c, g.z["cumsum_c"])
    A2 = A.copy()
    assert orc.cs_dupl(A2)
    g.check("Adupl", A2)


@pytest.mark.parametrize("name", ["bcsstk01", "bcsstk16"])
def test_make_sym(name):
    g = Golden(name)
    S = orc.make_sym(g.A())
    g.check("S", S)
    nnz, norm = KNOWN_SYM[name]
    assert S.nnz == nnz and abs(orc.cs_norm(S) - norm) <= 1e-2
    ST = orc.cs_transpose(S, True)
    g.check("ST", ST)
    g.check("SST", orc.cs_multiply(S, ST))


@pytest.mark.parametrize("name", FIXTURES)
def test_reference_known_answers(name):
    """The (m, n, nnz, 1-norm) table the reference's CSparseTest1 asserts."""
    _, (m, n, nnz, nrm, d), (nrmT, dT), (dnnz, dnrm, dd) = KNOWN[name]
    A = Golden(name).A()
    assert (A.m, A.n, A.nnz) == (m, n, nnz)
    assert abs(orc.cs_norm(A) - nrm) <= d
    AT = orc.cs_transpose(A, True)
    assert (AT.m, AT.n, AT.nnz) == (n, m, nnz)
    assert abs(orc.cs_norm(AT) - nrmT) <= dT
    C = orc.cs_multiply(A, AT)
    eye = orc.csc(m, m, np.arange(m + 1), np.arange(m), np.ones(m))
    D = orc.cs_add(C, eye, 1.0, orc.cs_norm(C))
    assert (D.m, D.n, D.nnz) == (m, m, dnnz)
    assert abs(orc.cs_norm(D) - dnrm) <= dd


def _mat(d):
    return orc.OMat(d["m"], d["n"], np.array(d["p"], np.int32), np.array(d["i"], np.int32),
                    None if d["x"] is None else np.array(d["x"], np.float64), d["nzmax"], d["nz"])


def _same(M, d):
    assert (M.m, M.n, M.nzmax) == (d["m"], d["n"], d["nzmax"])
    assert list(M.p) == d["p"] and list(M.i[: len(d["i"])]) == d["i"] and len(M.i) == len(d["i"])
    if d["x"] is None:
        assert M.x is None
    else:
        assert np.array_equal(np.array(d["x"]).view(np.int64), M.x.view(np.int64))


def test_edge_cases():
    e = edge_cases()
    E32 = orc.OMat(3, 2, np.zeros(3, np.int32), np.zeros(1, np.int32), np.zeros(1), 1)
    _same(orc.cs_transpose(E32, True), e["transpose_empty_3x2"])
    E23 = orc.OMat(2, 3, np.zeros(4, np.int32), np.zeros(1, np.int32), np.zeros(1), 1)
    _same(orc.cs_multiply(E23, E32), e["multiply_empty_2x3_3x2"])
    R = orc.csc(1, 2, [0, 1, 2], [0, 0], [1.0, -1.0])
    Cc = orc.csc(2, 1, [0, 2], [0, 1], [1.0, 1.0])
    _same(orc.cs_multiply(R, Cc), e["multiply_cancel"])
    T = orc.OMat(2, 2, np.zeros(2, np.int32), np.zeros(2, np.int32), np.zeros(2), 2, nz=1)
    assert orc.cs_transpose(T, True) is None and e["transpose_triplet_is_none"]
    assert orc.cs_multiply(T, T) is None and e["multiply_triplet_is_none"]
    assert orc.cs_gaxpy(T, np.ones(2), np.zeros(2)) is False and e["gaxpy_triplet_is_false"]
    assert orc.cs_gaxpy(R, None, np.zeros(1)) is False and e["gaxpy_none_x_is_false"]
    assert orc.cs_multiply(R, R) is None and e["multiply_dim_mismatch_is_none"]
    assert orc.cs_cumsum(None, np.ones(1, np.int32), 1) == e["cumsum_none"] == -1
    p = np.full(6, 9, np.int32)
    c = np.array([3, 0, 2, 5, 11], np.int32)
    assert orc.cs_cumsum(p, c, 4) == e["cumsum_t1_ret"]
    assert list(p) == e["cumsum_t1_p"] and list(c) == e["cumsum_t1_c"]
    p0 = np.array([5], np.int32)
    assert orc.cs_cumsum(p0, np.zeros(0, np.int32), 0) == e["cumsum_n0_ret"] and list(p0) == e["cumsum_n0_p"]
    Dm = _mat(e["transpose_dups_in"])
    DT = orc.cs_transpose(Dm, True)
    _same(DT, e["transpose_dups"])
    _same(orc.cs_multiply(Dm, DT), e["multiply_dups"])
    yd = np.array([0.5, -1.5, 2.5])
    orc.cs_gaxpy(Dm, np.array([2.0, -3.0]), yd)
    assert list(yd) == e["gaxpy_dups_y"]
    Sp = orc.csc(2, 2, [0, 2, 4], [0, 1, 0, 1], [-0.0, float("nan"), float("inf"), 5e-324])
    tsp = orc.cs_transpose(Sp, True)
    assert [int(v) for v in tsp.x.view(np.int64)] == e["transpose_special_x_bits"]
    assert list(tsp.i) == e["transpose_special_i"]
    Pn = orc.csc(3, 2, [0, 2, 3], [0, 2, 1], None)
    _same(orc.cs_transpose(Pn, True), e["transpose_pattern_only"])
    _same(orc.cs_multiply(Pn, orc.cs_transpose(Pn, False)), e["multiply_pattern_only"])
    Lg = orc.csc(2, 2, [0, 1, 2], [0, 1, 1, 0], [1.5, 2.5, 99.0, 98.0])
    _same(orc.cs_transpose(Lg, True), e["transpose_tail_ignored"])
