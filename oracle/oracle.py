"""ctypes front-end of the CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module (see csparse_oracle.c).
The product package ``csparse_cuda`` never does.

Every function mirrors one reference function (csparse.py, cited per function
in csparse_oracle.c) on numpy-backed matrices: ``p``/``i`` int32, ``x`` float64.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess
from dataclasses import dataclass
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_csparse.so")
_lib = None

_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, via oracle/Makefile)."""
    src = os.path.join(_HERE, "csparse_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_csparse.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_cumsum.restype = C.c_int64
        L.orc_cumsum.argtypes = [_i32p, _i32p, C.c_int32]
        L.orc_transpose.restype = C.c_int
        L.orc_transpose.argtypes = [C.c_int32, C.c_int32, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        L.orc_gaxpy.restype = C.c_int
        L.orc_gaxpy.argtypes = [C.c_int32, _i32p, _i32p, _f64p, _f64p, _f64p]
        mat = [C.c_int32, C.c_int32, _i32p, _i32p, _f64p]
        L.orc_multiply.restype = C.c_void_p
        L.orc_multiply.argtypes = mat + mat
        L.orc_add.restype = C.c_void_p
        L.orc_add.argtypes = mat + mat + [C.c_double, C.c_double]
        L.orc_result_free.argtypes = [C.c_void_p]
        L.orc_result_nnz.restype = C.c_int64
        L.orc_result_nnz.argtypes = [C.c_void_p]
        L.orc_result_nzmax.restype = C.c_int64
        L.orc_result_nzmax.argtypes = [C.c_void_p]
        L.orc_result_has_values.restype = C.c_int
        L.orc_result_has_values.argtypes = [C.c_void_p]
        L.orc_result_copy.argtypes = [C.c_void_p, _i32p, _i32p, _f64p]
        L.orc_norm.restype = C.c_double
        L.orc_norm.argtypes = [C.c_int32, _i32p, _f64p]
        L.orc_compress.restype = C.c_int
        L.orc_compress.argtypes = [C.c_int32, C.c_int32, C.c_int32, _i32p, _i32p, _f64p, _i32p, _i32p, _f64p]
        L.orc_dupl.restype = C.c_int64
        L.orc_dupl.argtypes = [C.c_int32, C.c_int32, _i32p, _i32p, _f64p]
        L.orc_fkeep.restype = C.c_int64
        L.orc_fkeep.argtypes = [C.c_int32, _i32p, _i32p, _f64p, C.c_int, C.c_double]
        L.orc_permute.restype = C.c_int
        L.orc_permute.argtypes = [C.c_int32, _i32p, _i32p, _f64p, _i32p, _i32p, _i32p, _i32p, _f64p]
        L.orc_symperm.restype = C.c_int64
        L.orc_symperm.argtypes = [C.c_int32, _i32p, _i32p, _f64p, _i32p, _i32p, _i32p, _f64p]
        _lib = L
    return _lib


def _ip(a):
    return a.ctypes.data_as(_i32p) if a is not None else None


def _fp(a):
    return a.ctypes.data_as(_f64p) if a is not None else None


@dataclass
class OMat:
    """numpy-backed mirror of the reference ``cs`` object (csparse.py:37-54)."""
    m: int
    n: int
    p: np.ndarray                 # int32, n+1 (CSC) or nz column indices (triplet)
    i: np.ndarray                 # int32
    x: Optional[np.ndarray]       # float64 or None (pattern only)
    nzmax: int = 0
    nz: int = -1                  # -1 => CSC

    @property
    def nnz(self) -> int:
        return int(self.p[self.n]) if self.nz < 0 else self.nz

    def copy(self) -> "OMat":
        return OMat(self.m, self.n, self.p.copy(), self.i.copy(),
                    None if self.x is None else self.x.copy(), self.nzmax, self.nz)


def csc(m, n, p, i, x=None) -> OMat:
    p = np.ascontiguousarray(p, dtype=np.int32)
    i = np.ascontiguousarray(i, dtype=np.int32)
    x = None if x is None else np.ascontiguousarray(x, dtype=np.float64)
    return OMat(int(m), int(n), p, i, x, nzmax=max(len(i), 1), nz=-1)


def cs_cumsum(p: np.ndarray, c: np.ndarray, n: int) -> int:
    if p is None or c is None:
        return -1
    assert p.dtype == np.int32 and c.dtype == np.int32
    return int(lib().orc_cumsum(_ip(p), _ip(c), n))


def cs_transpose(A: OMat, values=True) -> Optional[OMat]:
    if A is None or A.nz != -1:
        return None
    nnz = A.nnz
    nzmax = max(nnz, 1)
    has_x = bool(values) and A.x is not None
    Cp = np.zeros(A.m + 1, np.int32)
    Ci = np.zeros(nzmax, np.int32)
    Cx = np.zeros(nzmax, np.float64) if has_x else None
    rc = lib().orc_transpose(A.m, A.n, _ip(A.p), _ip(A.i), _fp(A.x) if has_x else None,
                             _ip(Cp), _ip(Ci), _fp(Cx))
    assert rc == 0
    return OMat(A.n, A.m, Cp, Ci, Cx, nzmax=nzmax, nz=-1)


def cs_gaxpy(A: OMat, x: np.ndarray, y: np.ndarray) -> bool:
    if A is None or A.nz != -1 or x is None or y is None:
        return False
    if A.x is None:
        raise TypeError("cs_gaxpy on a pattern-only matrix")
    assert x.dtype == np.float64 and y.dtype == np.float64
    rc = lib().orc_gaxpy(A.n, _ip(A.p), _ip(A.i), _fp(A.x), _fp(x), _fp(y))
    assert rc == 0
    return True


def _take_result(h, m, n) -> OMat:
    L = lib()
    nnz = int(L.orc_result_nnz(h))
    nzmax = int(L.orc_result_nzmax(h))
    has_x = bool(L.orc_result_has_values(h))
    Cp = np.zeros(n + 1, np.int32)
    Ci = np.zeros(nzmax, np.int32)
    Cx = np.zeros(nzmax, np.float64) if has_x else None
    L.orc_result_copy(h, _ip(Cp), _ip(Ci), _fp(Cx))
    L.orc_result_free(h)
    return OMat(m, n, Cp, Ci, Cx, nzmax=nzmax, nz=-1)


def cs_multiply(A: OMat, B: OMat) -> Optional[OMat]:
    if A is None or A.nz != -1 or B is None or B.nz != -1:
        return None
    if A.n != B.m:
        return None
    h = lib().orc_multiply(A.m, A.n, _ip(A.p), _ip(A.i), _fp(A.x),
                           B.m, B.n, _ip(B.p), _ip(B.i), _fp(B.x))
    if not h:
        raise MemoryError("orc_multiply")
    return _take_result(h, A.m, B.n)


def cs_add(A: OMat, B: OMat, alpha: float, beta: float) -> Optional[OMat]:
    if A is None or A.nz != -1 or B is None or B.nz != -1:
        return None
    if A.m != B.m or A.n != B.n:
        return None
    h = lib().orc_add(A.m, A.n, _ip(A.p), _ip(A.i), _fp(A.x),
                      B.m, B.n, _ip(B.p), _ip(B.i), _fp(B.x), float(alpha), float(beta))
    if not h:
        raise MemoryError("orc_add")
    return _take_result(h, A.m, A.n)


def cs_norm(A: OMat) -> float:
    if A is None or A.nz != -1 or A.x is None:
        return -1
    return float(lib().orc_norm(A.n, _ip(A.p), _fp(A.x)))


def cs_compress(m, n, Ti, Tj, Tx) -> OMat:
    Ti = np.ascontiguousarray(Ti, np.int32)
    Tj = np.ascontiguousarray(Tj, np.int32)
    Tx = None if Tx is None else np.ascontiguousarray(Tx, np.float64)
    nz = len(Ti)
    nzmax = max(nz, 1)
    Cp = np.zeros(n + 1, np.int32)
    Ci = np.zeros(nzmax, np.int32)
    Cx = np.zeros(nzmax, np.float64) if Tx is not None else None
    rc = lib().orc_compress(m, n, nz, _ip(Ti), _ip(Tj), _fp(Tx), _ip(Cp), _ip(Ci), _fp(Cx))
    assert rc == 0
    return OMat(m, n, Cp, Ci, Cx, nzmax=nzmax, nz=-1)


def cs_dupl(A: OMat) -> bool:
    if A is None or A.nz != -1:
        return False
    nz = int(lib().orc_dupl(A.m, A.n, _ip(A.p), _ip(A.i), _fp(A.x)))
    A.i = A.i[:nz].copy()
    A.x = A.x[:nz].copy()
    A.nzmax = nz
    return True


def cs_fkeep(A: OMat, mode: str, tol: float = 0.0) -> int:
    """mode in {'nonzero', 'tol', 'dropdiag'} (csparse.py:1002-1030; csparse_test.py Dropdiag)."""
    if A is None or A.nz != -1:
        return -1
    code = {"nonzero": 0, "tol": 1, "dropdiag": 2}[mode]
    nz = int(lib().orc_fkeep(A.n, _ip(A.p), _ip(A.i), _fp(A.x), code, float(tol)))
    A.i = A.i[:nz].copy()
    if A.x is not None:
        A.x = A.x[:nz].copy()
    A.nzmax = nz
    return nz


def cs_permute(A: OMat, pinv, q, values=True) -> Optional[OMat]:
    """C = P A Q (csparse.py:1666-1693)."""
    if A is None or A.nz != -1:
        return None
    nnz = A.nnz
    nzmax = max(nnz, 1)
    pinv = None if pinv is None else np.ascontiguousarray(pinv, np.int32)
    q = None if q is None else np.ascontiguousarray(q, np.int32)
    Cp = np.zeros(A.n + 1, np.int32)
    Ci = np.zeros(nzmax, np.int32)
    Cx = np.zeros(nzmax, np.float64) if (values and A.x is not None) else None
    rc = lib().orc_permute(A.n, _ip(A.p), _ip(A.i), _fp(A.x if Cx is not None else None),
                           _ip(pinv), _ip(q), _ip(Cp), _ip(Ci), _fp(Cx))
    assert rc == 0
    return OMat(A.m, A.n, Cp, Ci, Cx, nzmax=nzmax, nz=-1)


def cs_symperm(A: OMat, pinv, values=True) -> Optional[OMat]:
    """C = P A P', upper triangular part of a symmetric A (csparse.py:2220-2255)."""
    if A is None or A.nz != -1:
        return None
    nzmax = max(A.nnz, 1)
    pinv = None if pinv is None else np.ascontiguousarray(pinv, np.int32)
    Cp = np.zeros(A.n + 1, np.int32)
    Ci = np.zeros(nzmax, np.int32)
    Cx = np.zeros(nzmax, np.float64) if (values and A.x is not None) else None
    tot = lib().orc_symperm(A.n, _ip(A.p), _ip(A.i), _fp(A.x if Cx is not None else None),
                            _ip(pinv), _ip(Cp), _ip(Ci), _fp(Cx))
    assert tot >= 0
    return OMat(A.n, A.n, Cp, Ci, Cx, nzmax=nzmax, nz=-1)


def cs_amd_matrix(order: int, A: OMat) -> Optional[OMat]:
    """The matrix cs_amd builds before its elimination loop (csparse.py:228-258): pattern of
    A+A' (order 1, square), A'A without dense rows (order 2) or A'A (order 3), diagonal dropped."""
    if A is None or A.nz != -1 or order <= 0 or order > 3:
        return None
    AT = cs_transpose(A, False)
    m, n = A.m, A.n
    dense = min(n - 2, max(16, 10 * int(np.sqrt(n))))           # :233-234
    if order == 1 and n == m:
        C = cs_add(A, AT, 0, 0)                                   # :236
    elif order == 2:
        lens = np.diff(AT.p[: m + 1])                             # :238-249: drop dense columns of AT
        keepcol = lens <= dense
        keep = np.repeat(keepcol, lens)
        p2 = np.zeros(m + 1, np.int32)
        np.cumsum(np.where(keepcol, lens, 0), out=p2[1:])
        i2 = AT.i[: AT.nnz][keep]
        AT = OMat(AT.m, AT.n, p2, np.ascontiguousarray(i2 if len(i2) else np.zeros(1, np.int32)), None,
                  nzmax=max(len(i2), 1), nz=-1)
        A2 = cs_transpose(AT, False)
        C = cs_multiply(AT, A2)
    else:
        C = cs_multiply(AT, A)
    if C.nnz == 0 and len(C.i) == 0:
        C.i = np.zeros(1, np.int32)
    cs_fkeep(C, "dropdiag")                                        # :257
    return C


def make_sym(A: OMat) -> OMat:
    """C = A + triu(A,1)' as in csparse_test.py:115-121."""
    AT = cs_transpose(A, True)
    cs_fkeep(AT, "dropdiag")
    return cs_add(A, AT, 1.0, 1.0)


# ---- comparison helpers shared by the tests -------------------------------

def canonical(M: OMat) -> OMat:
    """Stable per-column sort by row index (parity form for cs_multiply)."""
    nnz = M.nnz
    cols = np.repeat(np.arange(M.n, dtype=np.int64), np.diff(M.p[: M.n + 1]).astype(np.int64))
    order = np.lexsort((M.i[:nnz], cols))
    return OMat(M.m, M.n, M.p.copy(), M.i[:nnz][order].copy(),
                None if M.x is None else M.x[:nnz][order].copy(), nzmax=M.nzmax, nz=-1)


def digest(*arrays) -> str:
    h = hashlib.sha256()
    for a in arrays:
        if a is None:
            h.update(b"<none>")
        else:
            a = np.ascontiguousarray(a)
            h.update(str(a.dtype).encode() + b":" + str(a.shape).encode() + b":")
            h.update(a.tobytes())
    return h.hexdigest()
