#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED Python reference.

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

For every matrix/ fixture of the reference it runs the reference's own
functions (csparse.py: cs_load :1307, cs_compress :647, cs_transpose :2292,
cs_multiply :1608, cs_gaxpy :1199, cs_cumsum :767, cs_add :163, cs_norm :1647,
cs_dupl :1035, cs_fkeep :1172) exactly as csparse_test.py's CSparseTest1 flow
does (csparse_test.py:235-266) and records inputs and outputs.  Small results
are stored in full; large ones (bcsstk16, mbeacxc) as sha256 digests of the
int32/float64 byte images plus sizes and 1-norms.  TEST INFRASTRUCTURE.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

REF = os.environ.get("CSPARSE_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import csparse as ref  # noqa: E402  (the unmodified reference)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle.oracle import digest  # noqa: E402
from csparse_cuda import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
FULL_LIMIT = 4000  # store arrays in full when nnz <= this

FIXTURES = ["t1", "ash219", "bcsstk01", "bcsstk16", "fs_183_1", "ibm32a", "ibm32b",
            "lp_afiro", "mbeacxc", "west0067"]


class Dropdiag(ref.cs_ifkeep):
    """csparse_test.py's Dropdiag predicate (keep off-diagonal entries)."""
    def fkeep(self, i, j, aij, other):
        return i != j


def nnz_of(A):
    return A.p[A.n]


def arrays(A):
    """(p, i, x) of a reference CSC object as numpy, logical prefix only."""
    nz = nnz_of(A)
    p = np.array(A.p[: A.n + 1], dtype=np.int32)
    i = np.array(A.i[:nz], dtype=np.int32)
    x = None if A.x is None else np.array(A.x[:nz], dtype=np.float64)
    return p, i, x


def canonical(p, i, x):
    n = len(p) - 1
    cols = np.repeat(np.arange(n, dtype=np.int64), np.diff(p).astype(np.int64))
    order = np.lexsort((i, cols))
    return p, i[order], None if x is None else x[order]


def record(store, meta, key, A, full):
    p, i, x = arrays(A)
    meta[key] = {
        "m": A.m, "n": A.n, "nnz": int(nnz_of(A)), "nzmax": int(A.nzmax),
        "len_i": len(A.i), "len_x": None if A.x is None else len(A.x),
        "has_x": A.x is not None,
        "norm1": ref.cs_norm(A) if A.x is not None else None,
        "sha": digest(p, i, x),
    }
    cp, ci, cx = canonical(p, i, x)
    meta[key]["sha_canonical_pattern"] = digest(cp, ci)
    meta[key]["sha_canonical"] = digest(cp, ci, cx)
    if full:
        store[key + "_p"] = p
        store[key + "_i"] = i
        if x is not None:
            store[key + "_x"] = x


def from_numpy(m, n, p, i, x):
    A = ref.cs()
    A.m, A.n, A.nz = m, n, -1
    A.p = [int(v) for v in p]
    A.i = [int(v) for v in i]
    A.x = None if x is None else [float(v) for v in x]
    A.nzmax = max(len(A.i), 1)
    return A


def run_flow(name, A, store, meta, extra_sym=False):
    """CSparseTest1 flow (csparse_test.py:251-266) + gaxpy + cumsum on a CSC A."""
    nnz = nnz_of(A)
    full = nnz <= FULL_LIMIT
    record(store, meta, "A", A, True)           # inputs always in full
    AT = ref.cs_transpose(A, True)
    record(store, meta, "AT", AT, full)
    ATp = ref.cs_transpose(A, False)
    record(store, meta, "ATpattern", ATp, False)
    ATT = ref.cs_transpose(AT, True)
    record(store, meta, "ATT", ATT, False)
    C = ref.cs_multiply(A, AT)
    record(store, meta, "C", C, full)
    Cpat = ref.cs_multiply(ATp, A)            # pattern-only product A'A (cs_amd style, :251)
    record(store, meta, "CpatternATA", Cpat, False)
    # D = C + norm(C) * I  (csparse_test.py:257-266)
    m = A.m
    T = ref.cs_spalloc(m, m, m, True, True)
    for k in range(m):
        ref.cs_entry(T, k, k, 1)
    Eye = ref.cs_compress(T)
    D = ref.cs_add(C, Eye, 1, ref.cs_norm(C))
    record(store, meta, "D", D, False)
    # gaxpy on A and AT
    x, y0 = synth.vectors(A.m, A.n)
    y = [float(v) for v in y0]
    assert ref.cs_gaxpy(A, [float(v) for v in x], y) is True
    store["gaxpy_y"] = np.array(y, dtype=np.float64)
    xt, yt0 = synth.vectors(A.n, A.m)
    yt = [float(v) for v in yt0]
    assert ref.cs_gaxpy(AT, [float(v) for v in xt], yt) is True
    store["gaxpy_yT"] = np.array(yt, dtype=np.float64)
    # cumsum on the row counts (what cs_transpose feeds it, :2307)
    cnt = np.bincount(np.array(A.i[:nnz], dtype=np.int64), minlength=A.m).astype(np.int32)
    c = [int(v) for v in cnt]
    pp = [7] * (A.m + 1)
    total = ref.cs_cumsum(pp, c, A.m)
    store["cumsum_in"] = cnt
    store["cumsum_p"] = np.array(pp, dtype=np.int32)
    store["cumsum_c"] = np.array(c, dtype=np.int32)
    meta["cumsum_total"] = int(total)
    # dupl (next row): on a copy
    A2 = from_numpy(A.m, A.n, *arrays(A))
    assert ref.cs_dupl(A2)
    record(store, meta, "Adupl", A2, False)
    if extra_sym:
        # make_sym as csparse_test.py:115-121
        AT2 = ref.cs_transpose(A, True)
        ref.cs_fkeep(AT2, Dropdiag(), None)
        S = ref.cs_add(A, AT2, 1, 1)
        record(store, meta, "S", S, False)
        ST = ref.cs_transpose(S, True)
        record(store, meta, "ST", ST, False)
        SST = ref.cs_multiply(S, ST)
        record(store, meta, "SST", SST, False)


def main():
    os.makedirs(OUT, exist_ok=True)
    index = {}
    for name in FIXTURES:
        store, meta = {}, {}
        T = ref.cs_load(os.path.join(REF, "matrix", name))
        store["T_i"] = np.array(T.i[: T.nz], dtype=np.int32)
        store["T_j"] = np.array(T.p[: T.nz], dtype=np.int32)
        store["T_x"] = np.array(T.x[: T.nz], dtype=np.float64)
        meta["T"] = {"m": T.m, "n": T.n, "nz": T.nz, "nzmax": T.nzmax}
        A = ref.cs_compress(T)
        run_flow(name, A, store, meta, extra_sym=name in ("bcsstk01", "bcsstk16"))
        store["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
        index[name] = {k: (v if not isinstance(v, dict) else
                           {kk: v[kk] for kk in ("m", "n", "nnz", "norm1") if kk in v})
                       for k, v in meta.items()}
        print(name, {k: (v["nnz"] if isinstance(v, dict) and "nnz" in v else v) for k, v in meta.items()})

    # synthetic families at oracle-friendly sizes
    for name, gen in (("lap2d_24", lambda: synth.lap2d(24)),
                      ("st27_7", lambda: synth.st27(7)),
                      ("rmat_9", lambda: synth.rmat(9, 8))):
        store, meta = {}, {}
        m, n, p, i, x = gen()
        A = from_numpy(m, n, p, i, x)
        run_flow(name, A, store, meta)
        if name.startswith("st27"):
            AA = ref.cs_multiply(A, A)
            record(store, meta, "AA", AA, False)
        store["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **store)
        print(name, {k: (v["nnz"] if isinstance(v, dict) and "nnz" in v else v) for k, v in meta.items()})

    # edge cases, all straight from the reference
    edge = {}

    def mk(m, n, p, i, x):
        return from_numpy(m, n, np.array(p), np.array(i), None if x is None else np.array(x))

    def dump(A):
        if A is None:
            return None
        return {"m": A.m, "n": A.n, "nz": A.nz, "nzmax": A.nzmax, "p": list(A.p), "i": list(A.i),
                "x": None if A.x is None else list(A.x)}

    E32 = mk(3, 2, [0, 0, 0], [], [])
    E32.nzmax = 1; E32.i = [0]; E32.x = [0.0]
    edge["transpose_empty_3x2"] = dump(ref.cs_transpose(E32, True))
    E23 = mk(2, 3, [0, 0, 0, 0], [], [])
    edge["multiply_empty_2x3_3x2"] = dump(ref.cs_multiply(E23, E32))
    R = mk(1, 2, [0, 1, 2], [0, 0], [1.0, -1.0])
    Cc = mk(2, 1, [0, 2], [0, 1], [1.0, 1.0])
    edge["multiply_cancel"] = dump(ref.cs_multiply(R, Cc))
    # triplet inputs -> sentinels
    Tt = ref.cs_spalloc(2, 2, 2, True, True)
    ref.cs_entry(Tt, 0, 0, 1.0)
    edge["transpose_triplet_is_none"] = ref.cs_transpose(Tt, True) is None
    edge["multiply_triplet_is_none"] = ref.cs_multiply(Tt, Tt) is None
    edge["gaxpy_triplet_is_false"] = ref.cs_gaxpy(Tt, [1.0, 1.0], [0.0, 0.0]) is False
    edge["gaxpy_none_x_is_false"] = ref.cs_gaxpy(R, None, [0.0]) is False
    edge["multiply_dim_mismatch_is_none"] = ref.cs_multiply(R, R) is None
    edge["cumsum_none"] = ref.cs_cumsum(None, [1], 1)
    pp = [9, 9, 9, 9, 9, 9]
    cc = [3, 0, 2, 5, 11]
    edge["cumsum_t1_ret"] = ref.cs_cumsum(pp, cc, 4)
    edge["cumsum_t1_p"] = pp
    edge["cumsum_t1_c"] = cc
    p0 = [5]
    edge["cumsum_n0_ret"] = ref.cs_cumsum(p0, [], 0)
    edge["cumsum_n0_p"] = p0
    # duplicates inside a column + unsorted column (stability of transpose)
    Dm = mk(3, 2, [0, 4, 6], [2, 0, 2, 0, 1, 1], [1.0, 2.0, 3.0, 4.0, 5.0, 6.0])
    edge["transpose_dups_in"] = dump(Dm)
    edge["transpose_dups"] = dump(ref.cs_transpose(Dm, True))
    edge["multiply_dups"] = dump(ref.cs_multiply(Dm, ref.cs_transpose(Dm, True)))
    yd = [0.5, -1.5, 2.5]
    ref.cs_gaxpy(Dm, [2.0, -3.0], yd)
    edge["gaxpy_dups_y"] = yd
    # -0.0, nan, inf, denormal survive transpose bit-for-bit
    Sp = mk(2, 2, [0, 2, 4], [0, 1, 0, 1], [-0.0, float("nan"), float("inf"), 5e-324])
    tsp = ref.cs_transpose(Sp, True)
    edge["transpose_special_x_bits"] = [int(v) for v in np.array(tsp.x, dtype=np.float64).view(np.int64)]
    edge["transpose_special_i"] = list(tsp.i)
    # pattern-only A (x None)
    Pn = mk(3, 2, [0, 2, 3], [0, 2, 1], None)
    edge["transpose_pattern_only"] = dump(ref.cs_transpose(Pn, True))
    edge["multiply_pattern_only"] = dump(ref.cs_multiply(Pn, ref.cs_transpose(Pn, False)))
    # longer-than-needed inputs: nzmax > nnz tails ignored
    Lg = mk(2, 2, [0, 1, 2], [0, 1, 1, 0], [1.5, 2.5, 99.0, 98.0])
    edge["transpose_tail_ignored"] = dump(ref.cs_transpose(Lg, True))
    with open(os.path.join(OUT, "edge_cases.json"), "w") as f:
        json.dump(edge, f, indent=1, allow_nan=True)
    print("edge cases:", len(edge))


if __name__ == "__main__":
    main()
