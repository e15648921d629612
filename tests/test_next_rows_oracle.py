"""CPU: the oracle's "next" rows against golden vectors made from the unmodified
reference by oracle/make_golden_next.py (cs_permute, cs_symperm, cs_norm, cs_add, cs_dropzeros,
cs_droptol) -- pins the checker the GPU tests use."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from tests.golden_util import FIXTURES, GOLDEN, Golden

with open(os.path.join(GOLDEN, "next_rows.json")) as f:
    NEXT = json.load(f)


def perm(n, seed):
    return np.random.default_rng(seed).permutation(n).astype(np.int32)


def pinv_of(p):
    pinv = np.empty_like(p)
    pinv[p] = np.arange(len(p), dtype=np.int32)
    return pinv


def check(M, g, what):
    nnz = int(M.p[M.n])
    assert (M.m, M.n, nnz) == (g["m"], g["n"], g["nnz"]), what
    x = None if M.x is None else np.asarray(M.x[:nnz], np.float64)
    assert (x is not None) == g["has_x"], what
    assert orc.digest(np.asarray(M.p[: M.n + 1], np.int32), np.asarray(M.i[:nnz], np.int32), x) == g["sha"], what


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_next_rows(name):
    g = NEXT[name]
    A = Golden(name).A()
    pinv = pinv_of(perm(A.m, 11))
    q = perm(A.n, 12)
    assert orc.digest(pinv) == g["pinv_sha"]
    check(orc.cs_permute(A, pinv, q, True), g["permute"], "permute")
    check(orc.cs_permute(A, None, q, False), g["permute_pattern_q_only"], "permute pattern")
    if A.m == A.n:
        check(orc.cs_symperm(A, pinv, True), g["symperm"], "symperm")
        check(orc.cs_symperm(A, None, False), g["symperm_identity_pattern"], "symperm identity")
        check(orc.cs_add(A, orc.cs_transpose(A, True), 1.0, 3.0), g["add_AT"], "add A+3A'")
    assert orc.cs_norm(A) == g["norm"]
    check(orc.cs_add(A, A, 2.5, -0.75), g["add_same"], "add same")
    A1 = A.copy()
    assert orc.cs_fkeep(A1, "nonzero") == g["dropzeros_ret"]
    check(A1, g["dropzeros"], "dropzeros")
    A2 = A.copy()
    assert orc.cs_fkeep(A2, "tol", g["droptol_tol"]) == g["droptol_ret"]
    check(A2, g["droptol"], "droptol")


@pytest.mark.parametrize("name", FIXTURES)
def test_oracle_amd_matrix(name):
    """the matrix cs_amd starts from (csparse.py:228-258): A+A', A'A without dense rows, A'A"""
    A = Golden(name).A()
    for order in (1, 2, 3):
        check(orc.cs_amd_matrix(order, A), NEXT[name]["amd_matrix_%d" % order], "amd matrix order %d" % order)
    assert orc.cs_amd_matrix(0, A) is None and orc.cs_amd_matrix(4, A) is None
