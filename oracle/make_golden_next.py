#!/usr/bin/env python
"""Golden vectors for the "next" rows of SURVEY.md 8f, from the UNMODIFIED Python reference:
cs_permute (csparse.py:1666), cs_symperm (:2220), cs_pinv (:1696), cs_dropzeros (:1024),
cs_droptol (:1007), cs_norm (:1647), cs_add with general coefficients (:163) on the reference's
matrix/ fixtures.  Run in the build container only (needs /root/reference):

    python oracle/make_golden_next.py      ->  tests/golden/next_rows.json

Results are stored as sha256 digests of the int32 / float64 images (oracle.digest) plus sizes;
inputs are rebuilt in the tests from tests/golden/<fixture>.npz and the seeds below.
TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

REF = os.environ.get("CSPARSE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import csparse as ref  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.oracle import digest  # noqa: E402
from oracle.make_golden import FIXTURES, arrays, from_numpy, nnz_of  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "next_rows.json")


def perm(n, seed):
    return np.random.default_rng(seed).permutation(n).astype(np.int32)


def rec(A):
    p, i, x = arrays(A)
    return {"m": A.m, "n": A.n, "nnz": int(nnz_of(A)), "nzmax": int(A.nzmax), "len_i": len(A.i),
            "has_x": A.x is not None, "sha": digest(p, i, x)}


def amd_matrix(order, A):
    """csparse.py:228-258 line for line, with the reference's own functions"""
    from math import sqrt
    AT = ref.cs_transpose(A, False)
    m, n = A.m, A.n
    dense = max(16, 10 * int(sqrt(n)))
    dense = min(n - 2, dense)
    if order == 1 and n == m:
        C = ref.cs_add(A, AT, 0, 0)
    elif order == 2:
        ATp, ATi = AT.p, AT.i
        p2 = 0
        for j in range(m):
            p = ATp[j]
            ATp[j] = p2
            if ATp[j + 1] - p > dense:
                continue
            while p < ATp[j + 1]:
                ATi[p2] = ATi[p]
                p2 += 1
                p += 1
        ATp[m] = p2
        A2 = ref.cs_transpose(AT, False)
        C = ref.cs_multiply(AT, A2)
    else:
        C = ref.cs_multiply(AT, A)
    ref.cs_fkeep(C, ref._cs_diag(), None)
    return C


def main():
    out = {}
    for name in FIXTURES:
        T = ref.cs_load(os.path.join(REF, "matrix", name))
        A = ref.cs_compress(T)
        g = {}
        p = perm(A.m, 11)                       # row permutation p, pinv = cs_pinv(p)
        q = perm(A.n, 12)
        pinv = ref.cs_pinv([int(v) for v in p], A.m)
        g["pinv_sha"] = digest(np.array(pinv, dtype=np.int32))
        g["permute"] = rec(ref.cs_permute(A, pinv, [int(v) for v in q], True))
        g["permute_pattern_q_only"] = rec(ref.cs_permute(A, None, [int(v) for v in q], False))
        if A.m == A.n:
            g["symperm"] = rec(ref.cs_symperm(A, pinv, True))
            g["symperm_identity_pattern"] = rec(ref.cs_symperm(A, None, False))
        g["norm"] = ref.cs_norm(A)
        # general add: C = 2.5*A - 0.75*A (same pattern) and A + A' for square fixtures
        g["add_same"] = rec(ref.cs_add(A, A, 2.5, -0.75))
        if A.m == A.n:
            g["add_AT"] = rec(ref.cs_add(A, ref.cs_transpose(A, True), 1.0, 3.0))
        # drop: on copies
        A1 = from_numpy(A.m, A.n, *arrays(A))
        g["dropzeros_ret"] = ref.cs_dropzeros(A1)
        g["dropzeros"] = rec(A1)
        tol = float(np.median(np.abs(np.array(A.x[:nnz_of(A)]))))
        A2 = from_numpy(A.m, A.n, *arrays(A))
        g["droptol_tol"] = tol
        g["droptol_ret"] = ref.cs_droptol(A2, tol)
        g["droptol"] = rec(A2)
        for order in (1, 2, 3):
            Cm = amd_matrix(order, A)
            r = rec(Cm)
            p_, i_, _x = arrays(Cm)
            n_ = len(p_) - 1
            cols = np.repeat(np.arange(n_, dtype=np.int64), np.diff(p_).astype(np.int64))
            r["sha_canonical_pattern"] = digest(p_, i_[np.lexsort((i_, cols))])
            g["amd_matrix_%d" % order] = r
        out[name] = g
        print(name, {k: (v["nnz"] if isinstance(v, dict) else v) for k, v in g.items()})
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
