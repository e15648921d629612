"""cs_transpose's one-pass mirror path (square matrix, strictly increasing columns, symmetric
pattern): taken when it applies, bit-identical to the reference (csparse.py:2292-2315) like every
other path, and backed out of -- with the general paths producing the same bits -- when a single
entry breaks what it relies on."""
import numpy as np
import pytest
import scipy.sparse as sp

import csparse_cuda as cc
from csparse_cuda import synth
from oracle import oracle as orc
from tests.test_gpu_parity import assert_same_matrix, bits, to_cs

pytestmark = pytest.mark.gpu


def sym_pattern(rng, n, density, dense_lines=0, values=True):
    """random square matrix with a symmetric pattern, sorted columns, UNsymmetric values"""
    S = sp.random(n, n, density=density, random_state=np.random.RandomState(int(rng.integers(1 << 30))), format="csr")
    S = S + S.T
    S.data[:] = 1.0
    S = S.tolil()
    for d in rng.integers(0, n, dense_lines):          # a dense row with its mirror column
        S[d, :] = 1.0
        S[:, d] = 1.0
    S = sp.csc_matrix(S)
    S.sort_indices()
    x = rng.standard_normal(S.nnz) if values else None
    return orc.csc(n, n, S.indptr.astype(np.int32), S.indices.astype(np.int32), x)


def check_all_paths(A, expect_auto, what):
    R = orc.cs_transpose(A, True)
    Rp = orc.cs_transpose(A, False)
    for path in (None, "bucket", "radix"):
        cc.force_transpose_path(path)
        try:
            C = cc.cs_transpose(to_cs(A, lists=False), True)
            took = cc.last_transpose_path()
            Cp = cc.cs_transpose(to_cs(A, lists=False), False)
        finally:
            cc.force_transpose_path(None)
        assert_same_matrix(C, R, f"{what} path {path}")
        assert_same_matrix(Cp, Rp, f"{what} pattern-only path {path}")
        if path is None and expect_auto == "general":
            assert took in ("bucket", "radix"), (what, took)
        elif path is None and expect_auto is not None:
            assert took == expect_auto, (what, took)
        if path == "radix":
            assert took == ("radix" if A.nnz else "trivial")
        if path == "bucket":
            assert took != "mirror"


@pytest.mark.parametrize("gen", [lambda: synth.lap2d(3), lambda: synth.lap2d(70), lambda: synth.lap2d(301),
                                 lambda: synth.st27(4), lambda: synth.st27(21)])
def test_mirror_taken_on_stencils(gen):
    m, n, p, i, x = gen()
    x = np.random.default_rng(5).standard_normal(len(x))      # unsymmetric values: a copy of A would be wrong
    check_all_paths(orc.csc(m, n, p, i, x), "mirror", "stencil")


@pytest.mark.parametrize("seed", range(12))
def test_mirror_random_symmetric_patterns(seed):
    """irregular columns (the positional guess misses, the binary search finds), dense lines,
    empty columns, pattern-only input"""
    rng = np.random.default_rng(300 + seed)
    n = int(rng.choice([1, 2, 7, 63, 400, 1500, 4000]))
    A = sym_pattern(rng, n, float(rng.choice([0.0, 0.002, 0.02, 0.2])), dense_lines=seed % 3, values=bool(seed % 4))
    check_all_paths(A, "mirror" if A.nnz else "trivial", f"random symmetric {seed}")


def test_mirror_many_empty_columns():
    """tiles spanning more columns than the staged pointer slice holds"""
    n = 60000
    rng = np.random.default_rng(11)
    r = np.sort(rng.choice(n, 600, replace=False))
    rows = np.concatenate([r, r[::-1]])                        # (r_k, r_{-k}) and its mirror
    cols = np.concatenate([r[::-1], r])
    S = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(n, n))
    S.data[:] = 1.0
    S.sort_indices()
    A = orc.csc(n, n, S.indptr.astype(np.int32), S.indices.astype(np.int32), rng.standard_normal(S.nnz))
    check_all_paths(A, "mirror", "sparse anti-diagonal")


def _lap(k):
    m, n, p, i, x = synth.lap2d(k)
    return m, n, p.copy(), i.copy(), np.random.default_rng(9).standard_normal(len(x))


def test_mirror_backs_out_on_one_missing_mirror():
    m, n, p, i, x = _lap(40)
    # drop the last entry of column 5: its mirror entry loses its partner
    q = p[6] - 1
    i2, x2 = np.delete(i, q), np.delete(x, q)
    p2 = p.copy(); p2[6:] -= 1
    A = orc.csc(m, n, p2, i2, x2)
    check_all_paths(A, "general", "one entry missing")
    # the handle remembers: a second transpose of the same device matrix skips the attempt
    dA = cc.upload(to_cs(A, lists=False))
    for _ in range(2):
        C = cc.cs_transpose(dA, True)
        assert cc.last_transpose_path() in ("bucket", "radix")
        assert_same_matrix(C.download(), orc.cs_transpose(A, True), "device handle")


def test_mirror_backs_out_on_unsorted_column():
    m, n, p, i, x = _lap(40)
    a, b = p[100], p[101]
    i[a:b] = i[a:b][::-1].copy()                               # same pattern, one column descending
    check_all_paths(orc.csc(m, n, p, i, x), "general", "unsorted column")


def test_mirror_backs_out_on_duplicate():
    m, n, p, i, x = _lap(40)
    q = p[300]                                                  # repeat the first entry of column 300
    i2, x2 = np.insert(i, q, i[q]), np.insert(x, q, 0.5)
    p2 = p.copy(); p2[301:] += 1
    check_all_paths(orc.csc(m, n, p2, i2, x2), "general", "duplicate entry")


def test_mirror_same_counts_but_not_symmetric():
    """row counts equal column counts, columns sorted, yet no entry has its mirror"""
    n = 500
    p = np.arange(n + 1, dtype=np.int32)
    i = ((np.arange(n) + 1) % n).astype(np.int32)              # a cyclic shift
    A = orc.csc(n, n, p, i, np.random.default_rng(2).standard_normal(n))
    check_all_paths(A, None, "cyclic shift")
    cc.cs_transpose(to_cs(A, lists=False), True)
    assert cc.last_transpose_path() != "mirror"


def test_mirror_result_feeds_gaxpy_and_round_trip():
    m, n, p, i, x = _lap(128)
    dA = cc.from_arrays(m, n, p, i, x)
    dC = cc.cs_transpose(dA, True)
    assert cc.last_transpose_path() == "mirror"
    p2, i2, x2 = cc.cs_transpose(dC, True).arrays()
    assert np.array_equal(p2, p) and np.array_equal(i2, i) and np.array_equal(bits(x2), bits(x))
    xv, y0 = synth.vectors(m, n)
    A = orc.csc(m, n, p, i, x)
    yr = y0.copy(); orc.cs_gaxpy(A, xv, yr)
    y = y0.copy(); assert cc.cs_gaxpy(to_cs(A, lists=False), xv, y)
    assert np.linalg.norm(y - yr) <= 1e-12 * np.linalg.norm(yr)
