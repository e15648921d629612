"""cs_multiply on pattern classes (csparse_cuda/csrc/spgemm_tpl.cuh) against the oracle.

Columns formed from a class template must equal the reference's cs_multiply (csparse.py:1608-1642)
bit for bit -- p, i in discovery order, x summed in the reference's sequence -- because the template
is cs_scatter's own trace on the class representative.  Matrices without translation-invariant
structure must fall through to the general kernels untouched.
"""
import numpy as np
import pytest

import csparse_cuda as cc
from csparse_cuda import synth
from oracle import oracle as orc
from tests.golden_util import ALL, Golden
from tests.test_gpu_parity import as_omat, assert_multiply_parity, assert_same_matrix, bits, to_cs

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["templates", "templates_percol"])
def templates(request):
    """both numeric kernels of the template path: 32 columns of a class in lock step on the
    entry-major copies (k_num_soa, the rest by k_num_tpl), and one warp per column only"""
    cc.force_multiply_path(request.param)
    yield request.param
    cc.force_multiply_path(None)


def _banded(n, offsets, seed=0, drop=()):
    """Toeplitz-like band matrix: column j holds rows j + d for d in offsets (inside the matrix),
    minus the entries listed in `drop` (breaks the invariance locally)."""
    rng = np.random.default_rng(seed)
    cols, rows = [], []
    for j in range(n):
        for d in offsets:
            i = j + d
            if 0 <= i < n and (i, j) not in drop:
                rows.append(i)
                cols.append(j)
    rows, cols = np.array(rows, np.int32), np.array(cols, np.int32)
    p = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(cols, minlength=n), out=p[1:])
    return n, n, p, rows, rng.uniform(-1.0, 1.0, len(rows))


@pytest.mark.parametrize("gen,all_templated", [
    (lambda: synth.lap2d(40), True),
    (lambda: synth.st27(9), True),
    (lambda: _banded(3000, (-7, -1, 0, 2, 40)), True),
    (lambda: _banded(3000, (5, 0, -3)), True),                               # unsorted columns are not canonical
    (lambda: _banded(2500, (-2, 0, 1), drop={(100, 100), (1501, 1500)}), True),
])
def test_templates_stencils_bit_exact(templates, gen, all_templated):
    m, n, p, i, x = gen()
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_multiply(A, A)
    dA = cc.from_arrays(m, n, p, i, x)
    dC = cc.cs_multiply(dA, dA)
    nt = cc.last_multiply_templated()
    canonical = bool(np.all([np.all(np.diff(i[p[j]:p[j + 1]]) > 0) for j in range(n)]))
    if canonical:
        assert nt == n, "every column of a translation-invariant operator has a template"
        assert_same_matrix(dC.download(trim=True), R, "templates: reference order, bit-exact")
    else:
        assert nt == 0                                         # duplicates / unsorted: general kernels
        assert_multiply_parity(dC.download(trim=True), R)
    # pattern only
    dP = cc.from_arrays(m, n, p, i, None)
    Cp = cc.cs_multiply(dP, dA).download(trim=True)
    assert Cp.x is None and np.array_equal(np.asarray(Cp.p), R.p)
    if canonical:
        assert np.array_equal(np.asarray(Cp.i), R.i[: R.nnz])


def test_templates_rectangular_column_block(templates):
    """B = a column block of A (what the column-sharded product multiplies): relative offsets shift."""
    m, n, p, i, x = synth.st27(8)
    A = orc.csc(m, n, p, i, x)
    dA = cc.from_arrays(m, n, p, i, x)
    for j0, j1 in ((0, 100), (137, 400), (n - 77, n), (50, 50)):
        b, e = int(p[j0]), int(p[j1])
        Bp = (p[j0:j1 + 1] - b).astype(np.int32)
        B = orc.csc(m, j1 - j0, Bp, i[b:e].copy(), x[b:e].copy())
        R = orc.cs_multiply(A, B)
        dB = dA.col_slice(j0, j1)
        C = cc.cs_multiply(dA, dB).download(trim=True)
        assert cc.last_multiply_templated() == (j1 - j0)
        assert_same_matrix(C, R, f"block [{j0},{j1})")


def test_templates_mixed_with_general_columns(templates):
    """A stencil with a few hundred perturbed columns: those fall to the general kernels, the rest
    keep their templates; the result is one consistent matrix."""
    m, n, p, i, x = synth.lap2d(48)
    import scipy.sparse as sp
    S = sp.csc_matrix((x, i, p), shape=(m, n)).tolil()
    rng = np.random.default_rng(4)
    for j in rng.choice(n, 40, replace=False):
        S[int(rng.integers(0, m)), int(j)] = 0.5
    S = S.tocsc(); S.sort_indices()
    A = orc.csc(m, n, S.indptr, S.indices, S.data)
    R = orc.cs_multiply(A, A)
    dA = cc.from_arrays(m, n, S.indptr, S.indices, S.data)
    C = cc.cs_multiply(dA, dA).download(trim=True)
    nt = cc.last_multiply_templated()
    assert 0 < nt <= n
    assert_multiply_parity(C, R, "mixed")
    cz, rz = orc.canonical(as_omat(C)), orc.canonical(R)
    assert np.array_equal(bits(cz.x), bits(rz.x))


@pytest.mark.parametrize("name", ALL)
def test_templates_forced_on_fixtures(templates, name):
    """Unstructured fixtures: one class per column (or a table overflow) -- same answers."""
    A = Golden(name).A()
    AT = orc.cs_transpose(A, True)
    R = orc.cs_multiply(A, AT)
    C = cc.cs_multiply(cc.upload(to_cs(A, False)), cc.upload(to_cs(AT, False))).download(trim=True)
    assert_multiply_parity(C, R, name)


def test_templates_rmat_overflows_the_class_table(templates):
    m, n, p, i, x = synth.rmat(12, 8)
    A = orc.csc(m, n, p, i, x)
    dA = cc.from_arrays(m, n, p, i, x)
    C = cc.cs_multiply(dA, dA).download(trim=True)
    assert cc.last_multiply_templated() == 0
    assert_multiply_parity(C, orc.cs_multiply(A, A), "rmat")


def test_templates_automatic_threshold_and_switch():
    m, n, p, i, x = synth.lap2d(130)                           # 16900 columns: above the automatic threshold
    A = orc.csc(m, n, p, i, x)
    R = orc.cs_multiply(A, A)
    dA = cc.from_arrays(m, n, p, i, x)
    C = cc.cs_multiply(dA, dA).download(trim=True)
    assert cc.last_multiply_templated() == n
    assert_same_matrix(C, R, "automatic")
    cc.force_multiply_path("no_templates")
    try:
        C2 = cc.cs_multiply(dA, dA).download(trim=True)
        assert cc.last_multiply_templated() == 0
    finally:
        cc.force_multiply_path(None)
    assert_multiply_parity(C2, R, "no templates")
    m, n, p, i, x = synth.lap2d(40)                            # 1600 columns: below it
    dS = cc.from_arrays(m, n, p, i, x)
    cc.cs_multiply(dS, dS)
    assert cc.last_multiply_templated() == 0


def test_templates_long_columns_fall_back(templates):
    """A class whose column would hold more than 256 rows or 4096 products has no template."""
    n = 4000
    offs = tuple(range(-20, 21))                              # 41 x 41 products = 1681, 81 rows: template
    m_, n_, p, i, x = _banded(n, offs)
    A = orc.csc(m_, n_, p, i, x)
    dA = cc.from_arrays(m_, n_, p, i, x)
    C = cc.cs_multiply(dA, dA).download(trim=True)
    assert cc.last_multiply_templated() == n
    assert_same_matrix(C, orc.cs_multiply(A, A), "wide band")
    offs = tuple(range(-70, 71))                              # 141 x 141 = 19881 products, 281 rows: no template
    m_, n_, p, i, x = _banded(n, offs, seed=2)
    A = orc.csc(m_, n_, p, i, x)
    dA = cc.from_arrays(m_, n_, p, i, x)
    C = cc.cs_multiply(dA, dA).download(trim=True)
    assert cc.last_multiply_templated() < n                    # boundary classes may still fit
    assert_multiply_parity(C, orc.cs_multiply(A, A), "very wide band")
