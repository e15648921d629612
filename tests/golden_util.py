"""Loaders for tests/golden/ (vectors produced by oracle/make_golden.py from the
unmodified reference) and the CSparseTest1 known-answer table."""
import json
import os

import numpy as np

from oracle import oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

FIXTURES = ["t1", "ash219", "bcsstk01", "bcsstk16", "fs_183_1", "ibm32a", "ibm32b",
            "lp_afiro", "mbeacxc", "west0067"]
SYNTH = ["lap2d_24", "st27_7", "rmat_9"]
ALL = FIXTURES + SYNTH

# Known answers asserted by the reference's own test-suite, CSparseTest1
# (csparse_test.py:269-426): name -> (line, A:(m, n, nnz, norm1, delta),
# AT norm1 (delta), D:(nnz, norm1, delta)).  D = A*A' + norm1(A*A')*I.
KNOWN = {
    "t1":       ("397-410", (4, 4, 10, 11.1, 1e-3), (7.7, 1e-3), (16, 139.58, 1e-3)),
    "ash219":   ("269-282", (219, 85, 438, 9, 1e-3), (2, 1e-3), (2205, 32, 1e-3)),
    "bcsstk01": ("285-298", (48, 48, 224, 3.00944e9, 1e4), (3.57095e9, 1e4), (764, 1.73403e19, 1e14)),
    "bcsstk16": ("301-314", (4884, 4884, 147631, 4.91422e9, 1e4), (5.47522e9, 1e4), (544856, 4.13336e19, 1e14)),
    "fs_183_1": ("317-330", (183, 183, 1069, 1.70318e9, 1e4), (8.22724e8, 1e3), (19665, 2.80249e18, 1e13)),
    "ibm32a":   ("333-346", (32, 31, 123, 7, 1e-3), (8, 1e-3), (386, 70, 1e-3)),
    "ibm32b":   ("349-362", (31, 32, 123, 8, 1e-3), (7, 1e-3), (373, 64, 1e-3)),
    "lp_afiro": ("365-378", (27, 51, 102, 3.429, 1e-3), (20.525, 1e-3), (153, 128.963, 1e-3)),
    "mbeacxc":  ("381-394", (492, 490, 49920, 0.928629, 1e-3), (16.5516, 1e-3), (157350, 19.6068, 1e-3)),
    "west0067": ("413-426", (67, 67, 299, 6.14337, 1e-3), (6.59006, 1e-3), (1041, 61.0906, 1e-3)),
}
# csparse_test.py:505,744 and :525,759 (make_sym results)
KNOWN_SYM = {"bcsstk01": (400, 3.5709480746974373e9), "bcsstk16": (290378, 7.008379365769155e9)}


class Golden:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.meta = json.loads(str(self.z["meta"]))

    def A(self):
        a = self.meta["A"]
        return orc.csc(a["m"], a["n"], self.z["A_p"], self.z["A_i"], self.z["A_x"])

    def has(self, key):
        return key + "_p" in self.z.files

    def mat(self, key):
        a = self.meta[key]
        x = self.z[key + "_x"] if key + "_x" in self.z.files else None
        return orc.csc(a["m"], a["n"], self.z[key + "_p"], self.z[key + "_i"], x)

    def check(self, key, M, order="exact"):
        """Compare an OMat-like result (m, n, p, i, x, nzmax) with the golden record."""
        g = self.meta[key]
        nnz = int(M.p[M.n])
        assert (M.m, M.n, nnz) == (g["m"], g["n"], g["nnz"]), (self.name, key)
        p = np.asarray(M.p[: M.n + 1], dtype=np.int32)
        i = np.asarray(M.i[:nnz], dtype=np.int32)
        x = None if M.x is None else np.asarray(M.x[:nnz], dtype=np.float64)
        assert (x is not None) == g["has_x"], (self.name, key, "x presence")
        if order == "exact":
            assert orc.digest(p, i, x) == g["sha"], (self.name, key, "p/i/x digest")
        else:
            c = orc.canonical(orc.csc(M.m, M.n, p, i, x))
            if order == "canonical":
                assert orc.digest(c.p, c.i, c.x) == g["sha_canonical"], (self.name, key)
            else:
                assert orc.digest(c.p, c.i) == g["sha_canonical_pattern"], (self.name, key)


def edge_cases():
    with open(os.path.join(GOLDEN, "edge_cases.json")) as f:
        return json.load(f)


def assert_values_close(x, xref, rtol=1e-12, scale=None, what=""):
    """|x - xref| <= rtol * scale elementwise; scale defaults to |xref| (bit-equal zeros ok)."""
    x = np.asarray(x, dtype=np.float64)
    xref = np.asarray(xref, dtype=np.float64)
    assert x.shape == xref.shape, what
    s = np.abs(xref) if scale is None else np.asarray(scale)
    err = np.abs(x - xref)
    bad = ~((err <= rtol * s) | ((x == xref)))
    # nan == nan never; treat matching nan patterns as equal
    bad &= ~(np.isnan(x) & np.isnan(xref))
    assert not bad.any(), f"{what}: {int(bad.sum())} entries off, max err {err[bad].max():.3e}"


def normwise(y, yref):
    d = np.linalg.norm(np.asarray(y) - np.asarray(yref))
    n = np.linalg.norm(np.asarray(yref))
    return d / n if n > 0 else d
