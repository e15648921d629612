"""csparse_cuda -- B200-native drop-in for CSparse.py's data-parallel kernels.

Same names, argument meaning and error behaviour as the reference module
(rwl/CSparse.py, file csparse.py):

    cs            csparse.py:37-54     matrix object (nzmax, m, n, p, i, x, nz)
    CS_CSC        csparse.py:113-119
    CS_TRIPLET    csparse.py:122-128
    cs_cumsum     csparse.py:767-784   p = cumsum(c), c <- p[0..n-1]
    cs_transpose  csparse.py:2292-2315 C = A'
    cs_gaxpy      csparse.py:1199-1213 y += A*x
    cs_multiply   csparse.py:1608-1642 C = A*B

and, built from the same kernels, the callers / data formats either side of that path:

    cs_add        csparse.py:163-192   C = alpha*A + beta*B
    cs_norm       csparse.py:1647-1663 1-norm
    cs_compress   csparse.py:647-672   triplet -> compressed column
    cs_dupl       csparse.py:1035-1063 sum duplicates
    cs_fkeep      csparse.py:1172-1196 (fixed predicates), cs_dropzeros :1024, cs_droptol :1007
    cs_permute    csparse.py:1666-1693 C = P A Q
    cs_symperm    csparse.py:2220-2255 C = P A P'
    cs_pinv       csparse.py:1696-1710 (host helper)

Every function runs on the GPU through libcsparse_b200.so (hand-written sm_100a
CUDA, include/csparse_b200.h) -- there is no CPU fallback.  ``cs`` objects may be
backed by Python lists (as in the reference), ``array.array`` or numpy arrays;
results come back as plain lists, exactly shaped like the reference's.

For repeated use keep matrices on the device: ``upload(A)`` returns a
``DeviceMatrix`` that every function above accepts in place of a ``cs``; results
of cs_transpose / cs_multiply on device matrices stay on the device
(``.download()`` turns them into a ``cs``).
"""
from __future__ import annotations

import array as _array
import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from ._lib import CSparseCudaError  # noqa: F401

__all__ = ["cs", "CS_CSC", "CS_TRIPLET", "cs_cumsum", "cs_transpose", "cs_gaxpy", "cs_multiply",
           "cs_add", "cs_norm", "cs_compress", "cs_dupl", "cs_fkeep", "cs_dropzeros", "cs_droptol",
           "cs_permute", "cs_symperm", "cs_pinv", "cs_amd_matrix", "cs_ifkeep", "KEEP_NONZERO", "KEEP_TOL",
           "KEEP_OFFDIAG", "KEEP_UPPER", "KEEP_SHORTCOL", "DeviceMatrix", "upload", "CSparseCudaError"]


class cs(object):
    """Matrix in compressed-column or triplet form (csparse.py:37-54)."""

    def __init__(self):
        self.nzmax = 0   # maximum number of entries
        self.m = 0       # number of rows
        self.n = 0       # number of columns
        self.p = []      # column pointers (size n+1) or col indices (size nzmax)
        self.i = []      # row indices, size nzmax
        self.x = []      # numerical values, size nzmax (None: pattern only)
        self.nz = 0      # # of entries in triplet matrix, -1 for compressed-col


def CS_CSC(A):
    """True if A is in column-compressed form (csparse.py:113-119)."""
    return A is not None and A.nz == -1


def CS_TRIPLET(A):
    """True if A is in triplet form (csparse.py:122-128)."""
    return A is not None and A.nz >= 0


# ---- marshalling ------------------------------------------------------------

def _i32(seq, count) -> np.ndarray:
    """First `count` entries of a list / array / ndarray as contiguous int32."""
    if isinstance(seq, np.ndarray):
        a = seq[:count]
    else:
        part = seq[:count] if len(seq) != count else seq
        if isinstance(part, list):
            # a list of Python ints (the reference's own containers): array('i') converts and
            # range-checks in one pass, 1.5x faster than np.asarray + min/max on 1e7 entries
            try:
                return np.frombuffer(_array.array("i", part), dtype=np.int32) if part else np.zeros(0, np.int32)
            except (TypeError, OverflowError):
                pass                      # floats, numpy scalars without __index__, values past int32: slow path
        a = np.asarray(part)
    if a.size and a.dtype != np.int32:
        if a.dtype.kind not in "iu":
            a = a.astype(np.int64)
        if a.size and (a.max() > 0x7FFFFFFF or a.min() < -0x80000000):
            raise OverflowError("index does not fit int32")
    return np.ascontiguousarray(a, dtype=np.int32)


def _f64(seq, count) -> np.ndarray:
    if isinstance(seq, np.ndarray):
        a = seq[:count]
    else:
        part = seq[:count] if len(seq) != count else seq
        if isinstance(part, list) and part:
            try:
                return np.frombuffer(_array.array("d", part), dtype=np.float64)
            except TypeError:
                pass                      # entries that are not real numbers: let numpy decide
        a = np.asarray(part, dtype=np.float64)
    return np.ascontiguousarray(a, dtype=np.float64)


def _need(seq, count, what):
    """The reference indexes seq[0..count-1] and raises IndexError on a short container; the C ABI
    gets raw pointers, so lengths are checked here, before any pointer crosses the boundary."""
    if count < 0 or len(seq) < count:
        raise IndexError(f"{what}: {len(seq)} entries, {count} needed")


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class DeviceMatrix(object):
    """A CSC matrix resident in B200 HBM (opaque csb200_mat handle).

    Duck-types the read-only part of ``cs``: ``m``, ``n``, ``nz == -1``,
    ``nzmax``; ``p``/``i``/``x`` live on the device (see ``download``).
    """

    nz = -1

    def __init__(self, handle, owner=True):
        if not handle:
            raise CSparseCudaError("DeviceMatrix: null csb200_mat handle (the call that should have made it failed)")
        self._h = C.c_void_p(handle)
        self._owner = owner
        m, n, nnz, hv = C.c_int32(), C.c_int32(), C.c_int64(), C.c_int()
        _lib.check(_lib.lib().csb200_mat_dims(self._h, C.byref(m), C.byref(n), C.byref(nnz), C.byref(hv)))
        self.m, self.n, self.nnz, self.has_values = m.value, n.value, nnz.value, bool(hv.value)
        self.nzmax = max(self.nnz, 1)

    def __del__(self):
        try:
            if self._owner and self._h:
                _lib.lib().csb200_mat_free(self._h)
                self._h = None
        except Exception:
            pass

    def free(self):
        if self._owner and self._h:
            _lib.lib().csb200_mat_free(self._h)
        self._h = None

    def arrays(self):
        """(p, i, x) as numpy arrays of the logical sizes n+1 / nnz / nnz (x may be None)."""
        p = np.empty(self.n + 1, np.int32)
        i = np.empty(self.nnz, np.int32)
        x = np.empty(self.nnz, np.float64) if self.has_values else None
        _lib.check(_lib.lib().csb200_mat_download(self._h, _ptr(p), _ptr(i), _ptr(x)), "download")
        return p, i, x

    def device_pointers(self):
        """Raw device addresses (p, i, x) for interop with torch / other CUDA code."""
        dp, di, dx = C.c_void_p(), C.c_void_p(), C.c_void_p()
        _lib.check(_lib.lib().csb200_mat_dev_ptrs(self._h, C.byref(dp), C.byref(di), C.byref(dx)))
        return dp.value, di.value, dx.value

    def col_slice(self, j0: int, j1: int) -> "DeviceMatrix":
        out = C.c_void_p()
        st = _lib.check(_lib.lib().csb200_mat_col_slice(self._h, j0, j1, C.byref(out)), "col_slice", allow_arg=True)
        if st == _lib.ERR_ARG:
            raise ValueError(_lib.last_error())
        return DeviceMatrix(out.value)

    def download(self, trim: bool = False, numpy: bool = False) -> cs:
        """Back to a host ``cs``: list-backed like the reference's, or numpy-backed (``numpy=True``:
        int32 p / i, float64 x, no per-element conversion -- what the module functions return
        when their operands were numpy-backed).  ``trim`` selects cs_multiply's shape convention
        (nzmax == nnz, possibly 0) instead of cs_spalloc's max(nnz, 1)."""
        p, i, x = self.arrays()
        A = cs()
        A.m, A.n, A.nz = self.m, self.n, -1
        A.p = p if numpy else p.tolist()
        A.i = i if numpy else i.tolist()
        A.x = None if x is None else (x if numpy else x.tolist())
        if trim:
            A.nzmax = self.nnz
        else:
            A.nzmax = max(self.nnz, 1)
            if self.nnz == 0:
                A.i = np.zeros(1, np.int32) if numpy else [0]
                A.x = None if x is None else (np.zeros(1, np.float64) if numpy else [0.0])
        return A

    # device-pointer forms (async on the stream set with set_stream); x / y are raw
    # device addresses, e.g. torch_tensor.data_ptr()
    def gaxpy_dev(self, x_ptr: int, y_ptr: int):
        """y += A*x on device vectors (csb200_gaxpy_dev)."""
        _lib.check(_lib.lib().csb200_gaxpy_dev(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr)), "gaxpy_dev")

    def gaxpy_t_dev(self, x_ptr: int, y_ptr: int):
        """y[0..n) += A'*x[0..m): this matrix used directly as a CSR view (csb200_gaxpy_t_dev)."""
        _lib.check(_lib.lib().csb200_gaxpy_t_dev(self._h, C.c_void_p(x_ptr), C.c_void_p(y_ptr)), "gaxpy_t_dev")

    # gaxpy plan control (tests / benchmarks)
    def prepare_gaxpy(self):
        _lib.check(_lib.lib().csb200_gaxpy_prepare(self._h), "gaxpy_prepare")

    def gaxpy_plan(self) -> str:
        k = C.c_int()
        _lib.check(_lib.lib().csb200_gaxpy_plan(self._h, C.byref(k)), "gaxpy_plan")
        return {1: "stream", 2: "merge", 3: "stream_ld", 4: "split"}.get(k.value, "none")

    def force_gaxpy_plan(self, kind: Optional[str]):
        code = {None: 0, "auto": 0, "stream": 1, "merge": 2, "stream_ld": 3, "split": 4}[kind]
        _lib.check(_lib.lib().csb200_gaxpy_force_plan(self._h, code))


def force_transpose_path(path: Optional[str]):
    """Tests / benchmarks: None or "auto" = automatic choice, "radix" = always the stable radix sort,
    "bucket" = automatic choice without the one-pass mirror path, "bucket_slab" = the same with the
    partition and the sort interleaved in L2-sized slabs (measured slower; A/B measurements),
    "bucket_fused" = partition and sort in one persistent launch, a bucket sorted by the CTA that
    completes it (measured slower as well)."""
    _lib.check(_lib.lib().csb200_transpose_force_path({None: 0, "auto": 0, "radix": 1, "bucket": 2,
                                                       "bucket_slab": 3, "bucket_fused": 4}[path]))


def last_transpose_path() -> str:
    """Which kernel path this thread's last cs_transpose took ("mirror", "bucket", "radix", "trivial")."""
    return {1: "mirror", 2: "bucket", 3: "radix"}.get(_lib.lib().csb200_transpose_last_path(), "trivial")


def upload(A, validate: bool = True) -> DeviceMatrix:
    """Copy a CSC ``cs`` (lists, array.array or numpy) into HBM."""
    if isinstance(A, DeviceMatrix):
        return A
    if not CS_CSC(A):
        raise ValueError("upload: a compressed-column cs is required")
    n = A.n
    if A.m < 0 or n < 0:
        raise ValueError("upload: negative dimension")
    _need(A.p, n + 1, "upload: column pointers p")
    p = _i32(A.p, n + 1)
    nnz = int(p[n])
    if nnz < 0 or nnz > len(A.i) or (A.x is not None and nnz > len(A.x)):
        raise ValueError("upload: p[n] exceeds the length of i / x")
    i = _i32(A.i, nnz)
    x = None if A.x is None else _f64(A.x, nnz)
    out = C.c_void_p()
    st = _lib.check(_lib.lib().csb200_mat_upload(A.m, n, _ptr(p), _ptr(i), _ptr(x), 1 if validate else 0,
                                                 C.byref(out)), "upload", allow_arg=True)
    if st == _lib.ERR_ARG:
        raise ValueError(_lib.last_error())
    return DeviceMatrix(out.value)


def from_arrays(m, n, p, i, x=None, validate: bool = True) -> DeviceMatrix:
    """Upload numpy CSC arrays without building a ``cs``."""
    A = cs()
    A.m, A.n, A.nz, A.p, A.i, A.x = m, n, -1, p, i, x
    A.nzmax = max(len(i), 1)
    return upload(A, validate)


def from_device(m, n, p_ptr: int, i_ptr: int, x_ptr: int = 0) -> DeviceMatrix:
    """New handle from CSC arrays already in device memory (raw addresses, e.g.
    torch_tensor.data_ptr(); int32 p of n+1 entries, int32 i, float64 x or 0 for
    pattern-only).  The arrays are copied; contents are trusted (not validated)."""
    out = C.c_void_p()
    st = _lib.check(_lib.lib().csb200_mat_from_dev(int(m), int(n), C.c_void_p(p_ptr), C.c_void_p(i_ptr),
                                                   C.c_void_p(x_ptr or None), C.byref(out)), "from_device", allow_arg=True)
    if st == _lib.ERR_ARG:
        raise ValueError(_lib.last_error())
    return DeviceMatrix(out.value)


def _numpy_backed(*ops) -> bool:
    """True when a host operand keeps its row indices in a numpy array: results then stay numpy
    arrays (a list-backed cs in gives a list-backed cs out, as the reference's callers expect)."""
    return any(isinstance(getattr(o, "i", None), np.ndarray) for o in ops if not isinstance(o, DeviceMatrix))


def _as_device(A):
    """(DeviceMatrix, temporary?) for a cs or DeviceMatrix operand."""
    if isinstance(A, DeviceMatrix):
        return A, False
    return upload(A), True


# ---- the four reference functions ---------------------------------------------

def cs_cumsum(p, c, n):
    """p [0..n] = cumulative sum of c [0..n-1], and then copy p [0..n-1] into c.

    Reference: csparse.py:767-784.  Returns sum(c), or -1 if p or c is None.
    Runs the decoupled look-back scan kernel (csb200_cumsum).
    """
    if p is None or c is None:
        return -1
    n = int(n)
    if n < 0:
        n = 0                      # range(n) is empty in the reference: only p[0] = 0 is written
    _need(c, n, "cs_cumsum: c")
    _need(p, n + 1, "cs_cumsum: p")
    cc = _i32(c, n).copy()
    pp = np.empty(n + 1, np.int32)
    total = C.c_int64()
    _lib.check(_lib.lib().csb200_cumsum(_ptr(pp), _ptr(cc), n, C.byref(total)), "cs_cumsum")
    if isinstance(p, np.ndarray):
        p[: n + 1] = pp
    else:
        p[: n + 1] = pp.tolist()
    if n > 0:
        if isinstance(c, np.ndarray):
            c[:n] = cc
        else:
            c[:n] = cc.tolist()
    return int(total.value)


def cs_transpose(A, values):
    """Computes the transpose of a sparse matrix, C = A'.

    Reference: csparse.py:2292-2315.  Returns None unless A is compressed-column.
    The result's p, i and x are bit-identical to the reference's (stable order
    inside every column).  A DeviceMatrix in gives a DeviceMatrix out.
    """
    if not CS_CSC(A):
        return None
    dA, tmp = _as_device(A)
    out = C.c_void_p()
    _lib.check(_lib.lib().csb200_transpose(dA._h, 1 if values else 0, C.byref(out)), "cs_transpose")
    dC = DeviceMatrix(out.value)
    if not tmp:
        return dC
    dA.free()
    return dC.download(trim=False, numpy=_numpy_backed(A))


def cs_gaxpy(A, x, y):
    """Sparse matrix times dense column vector, y = A*x+y.

    Reference: csparse.py:1199-1213.  Returns False if A is not compressed-column
    or x / y is None, True otherwise; y is updated in place.  A pattern-only
    matrix raises TypeError (the reference fails the same way on ``None * x``).
    """
    if not CS_CSC(A) or x is None or y is None:
        return False
    has_x = A.has_values if isinstance(A, DeviceMatrix) else A.x is not None
    if not has_x:
        raise TypeError("cs_gaxpy: matrix has no numerical values (A.x is None)")
    dA, tmp = _as_device(A)
    try:
        _need(x, dA.n, "cs_gaxpy: x")
        _need(y, dA.m, "cs_gaxpy: y")
    except IndexError:
        if tmp:
            dA.free()
        raise
    xx = _f64(x, dA.n)
    inplace = (isinstance(y, np.ndarray) and y.dtype == np.float64 and y.flags.c_contiguous
               and y.flags.writeable)
    ybuf = y[: dA.m] if inplace else np.array(y[: dA.m], dtype=np.float64)
    _lib.check(_lib.lib().csb200_gaxpy(dA._h, _ptr(xx), _ptr(ybuf)), "cs_gaxpy")
    if tmp:
        dA.free()
    if not inplace:
        y[: dA.m] = ybuf.tolist() if isinstance(y, list) else ybuf
    return True


def cs_multiply(A, B):
    """Sparse matrix multiplication, C = A*B.

    Reference: csparse.py:1608-1642 (inner kernel cs_scatter, :1961-1989).
    Returns None unless both are compressed-column and A.n == B.m.  Structural
    zeros are kept; C.x is None when A or B is pattern-only; nzmax == nnz(C)
    (possibly 0).  Host ``cs`` operands: the columns of C are in the reference's
    discovery order (p, i, x bit-identical on canonical inputs).  Two device
    matrices: the faster blocked numeric kernel may emit a column's rows block by
    block (same set of rows, same values; see ``force_multiply_path``).
    """
    if not CS_CSC(A) or not CS_CSC(B):
        return None
    if A.n != B.m:
        return None
    dA, ta = _as_device(A)
    dB, tb = (dA, False) if B is A else _as_device(B)
    out = C.c_void_p()
    both_dev = isinstance(A, DeviceMatrix) and isinstance(B, DeviceMatrix)
    fn = _lib.lib().csb200_multiply if both_dev else _lib.lib().csb200_multiply_ordered
    _lib.check(fn(dA._h, dB._h, C.byref(out)), "cs_multiply")
    dC = DeviceMatrix(out.value)
    if ta:
        dA.free()
    if tb:
        dB.free()
    if isinstance(A, DeviceMatrix) and isinstance(B, DeviceMatrix):
        return dC
    return dC.download(trim=True, numpy=_numpy_backed(A, B))


# ---- the callers and data formats either side of the hot path ----------

def _padded(dC: "DeviceMatrix", nzmax: int, numpy: bool = False) -> cs:
    """Download into a cs whose i / x have cs_spalloc's length nzmax (zero tail); lists, or numpy
    arrays when the operands were numpy-backed."""
    p, i, x = dC.arrays()
    A = cs()
    A.m, A.n, A.nz, A.nzmax = dC.m, dC.n, -1, nzmax
    pad = nzmax - len(i)
    if numpy:
        A.p = p
        A.i = np.concatenate([i, np.zeros(pad, np.int32)]) if pad else i
        A.x = None if x is None else (np.concatenate([x, np.zeros(pad, np.float64)]) if pad else x)
        return A
    A.p = p.tolist()
    A.i = i.tolist() + [0] * pad
    A.x = None if x is None else x.tolist() + [0.0] * pad
    return A


def _nnz_of(A) -> int:
    return A.nnz if isinstance(A, DeviceMatrix) else int(A.p[A.n])


def cs_add(A, B, alpha, beta):
    """C = alpha*A + beta*B.

    Reference: csparse.py:163-192.  None unless both are compressed-column with equal
    dimensions.  Runs as [A B] * [alpha I; beta I] on the SpGEMM kernels, so the columns of C
    are in the reference's order (A's entries, then B's new rows) with its rounding.
    C.nzmax = max(nnz(A) + nnz(B), 1) as cs_spalloc leaves it (the reference does not trim).
    """
    if not CS_CSC(A) or not CS_CSC(B):
        return None
    if A.m != B.m or A.n != B.n:
        return None
    nzmax = max(_nnz_of(A) + _nnz_of(B), 1)
    dA, ta = _as_device(A)
    dB, tb = (dA, False) if B is A else _as_device(B)
    out = C.c_void_p()
    _lib.check(_lib.lib().csb200_add(dA._h, dB._h, float(alpha), float(beta), C.byref(out)), "cs_add")
    dC = DeviceMatrix(out.value)
    if ta:
        dA.free()
    if tb:
        dB.free()
    if isinstance(A, DeviceMatrix) and isinstance(B, DeviceMatrix):
        return dC
    return _padded(dC, nzmax, _numpy_backed(A, B))


def force_add_path(path: Optional[str]):
    """Tests: None / "auto" = automatic, "spgemm" = cs_add always on the SpGEMM kernels."""
    _lib.check(_lib.lib().csb200_add_force_path({None: 0, "auto": 0, "spgemm": 1}[path]))


def cs_norm(A):
    """1-norm of a sparse matrix = largest column sum of |x| (csparse.py:1647-1663).
    -1 if A is not compressed-column or has no values."""
    if not CS_CSC(A):
        return -1
    if (not A.has_values) if isinstance(A, DeviceMatrix) else (A.x is None):
        return -1
    dA, tmp = _as_device(A)
    v = C.c_double()
    _lib.check(_lib.lib().csb200_norm(dA._h, C.byref(v)), "cs_norm")
    if tmp:
        dA.free()
    return v.value


def compress_device(T) -> "DeviceMatrix":
    """cs_compress leaving the result in HBM."""
    nz = T.nz
    _need(T.i, nz, "cs_compress: row indices i")
    _need(T.p, nz, "cs_compress: column indices p")
    if T.x is not None:
        _need(T.x, nz, "cs_compress: values x")
    ti, tj = _i32(T.i, nz), _i32(T.p, nz)
    tx = None if T.x is None else _f64(T.x, nz)
    out = C.c_void_p()
    st = _lib.check(_lib.lib().csb200_compress(T.m, T.n, nz, _ptr(ti), _ptr(tj), _ptr(tx), C.byref(out)),
                    "cs_compress", allow_arg=True)
    if st == _lib.ERR_ARG:
        raise ValueError(_lib.last_error())
    return DeviceMatrix(out.value)


def cs_compress(T):
    """C = compressed-column form of a triplet matrix T (csparse.py:647-672).  None unless T is
    a triplet matrix.  Columns are not sorted and duplicates stay; the entries of a column keep
    their input order (stable radix sort on the column index)."""
    if not CS_TRIPLET(T):
        return None
    return compress_device(T).download(trim=False, numpy=_numpy_backed(T))


def cs_dupl(A):
    """Removes and sums duplicate entries (csparse.py:1035-1063).  In place on a list-backed
    cs: p, i, x are replaced, nzmax = nnz; returns True (False unless compressed-column).
    Device matrices are immutable: use ``dupl_device``."""
    if not CS_CSC(A):
        return False
    if isinstance(A, DeviceMatrix):
        raise TypeError("cs_dupl works in place on a host cs; use dupl_device(A) for device matrices")
    if A.x is None:
        raise TypeError("cs_dupl: matrix has no numerical values (A.x is None)")
    dC = dupl_device(A)
    p, i, x = dC.arrays()
    A.p[: A.n + 1] = p.tolist()
    A.i, A.x, A.nzmax = i.tolist(), x.tolist(), dC.nnz
    return True


def dupl_device(A) -> "DeviceMatrix":
    dA, tmp = _as_device(A)
    out = C.c_void_p()
    st = _lib.check(_lib.lib().csb200_dupl(dA._h, C.byref(out)), "cs_dupl", allow_arg=True)
    if tmp:
        dA.free()
    if st == _lib.ERR_ARG:
        raise TypeError(_lib.last_error())
    return DeviceMatrix(out.value)


KEEP_NONZERO, KEEP_TOL, KEEP_OFFDIAG, KEEP_UPPER, KEEP_SHORTCOL = 0, 1, 2, 3, 4


class cs_ifkeep(object):
    """The reference's predicate interface (csparse.py:1160-1169).  The GPU path evaluates a
    fixed set of predicates; subclasses name one through ``code``."""
    code = None

    def fkeep(self, i, j, aij, other):
        raise NotImplementedError


class _cs_nonzero(cs_ifkeep):
    code = KEEP_NONZERO

    def fkeep(self, i, j, aij, other):
        return aij != 0


class _cs_tol(cs_ifkeep):
    code = KEEP_TOL

    def fkeep(self, i, j, aij, other):
        return abs(aij) > float(other)


class cs_offdiag(cs_ifkeep):
    """csparse_test.py's Dropdiag: keep the off-diagonal entries."""
    code = KEEP_OFFDIAG

    def fkeep(self, i, j, aij, other):
        return i != j


class cs_upper(cs_ifkeep):
    code = KEEP_UPPER

    def fkeep(self, i, j, aij, other):
        return i <= j


def fkeep_device(A, code: int, tol: float = 0.0) -> "DeviceMatrix":
    dA, tmp = _as_device(A)
    out = C.c_void_p()
    _lib.check(_lib.lib().csb200_fkeep(dA._h, int(code), float(tol), C.byref(out)), "cs_fkeep")
    if tmp:
        dA.free()
    return DeviceMatrix(out.value)


def cs_fkeep(A, fkeep, other):
    """Drops entries from a sparse matrix (csparse.py:1172-1196); in place, returns the new
    number of entries, -1 unless compressed-column.  ``fkeep`` is one of the predicate objects
    of this module (cs_dropzeros / cs_droptol's, cs_offdiag, cs_upper) or a KEEP_* code: the
    predicate runs on the GPU, arbitrary Python callables cannot."""
    if not CS_CSC(A):
        return -1
    code = fkeep if isinstance(fkeep, int) else getattr(fkeep, "code", None)
    if code is None:
        raise NotImplementedError("cs_fkeep on the GPU supports the fixed predicates KEEP_NONZERO, "
                                  "KEEP_TOL, KEEP_OFFDIAG, KEEP_UPPER")
    if isinstance(A, DeviceMatrix):
        raise TypeError("cs_fkeep works in place on a host cs; use fkeep_device(A, code, tol)")
    dC = fkeep_device(A, code, 0.0 if other is None else float(other))
    p, i, x = dC.arrays()
    A.p[: A.n + 1] = p.tolist()
    A.i = i.tolist()
    A.x = None if A.x is None else x.tolist()
    A.nzmax = dC.nnz
    return dC.nnz


def cs_droptol(A, tol):
    """Removes entries with absolute value <= tol (csparse.py:1007-1014)."""
    return cs_fkeep(A, _cs_tol(), tol)


def cs_dropzeros(A):
    """Removes numerically zero entries (csparse.py:1024-1030)."""
    return cs_fkeep(A, _cs_nonzero(), None)


def cs_amd_matrix(order, A):
    """The matrix cs_amd orders (csparse.py:228-258), pattern only, diagonal dropped:
    order 1 (and A square): A + A'; order 2: A'A without the dense rows of A (more than
    max(16, 10*int(sqrt(n))), capped at n-2, entries); order 3: A'A.  None for a non-CSC A or
    another order.  Built from cs_transpose, cs_fkeep, cs_add and cs_multiply on the GPU; the
    sequential elimination that follows in cs_amd stays out of scope.  A device matrix in gives a
    device matrix out (rows of a column in the blocked kernel's order); a host cs gives a host cs
    whose p and i equal the reference's C bit for bit."""
    if not CS_CSC(A) or order <= 0 or order > 3:
        return None
    host = not isinstance(A, DeviceMatrix)
    dA, tmp = _as_device(A)
    mult = (lambda X, Y: DeviceMatrix(_mult_ordered(X, Y))) if host else cs_multiply
    dAT = cs_transpose(dA, False)
    m, n = dA.m, dA.n
    dense = min(n - 2, max(16, 10 * int(np.sqrt(n))))
    if order == 1 and n == m:
        dC = cs_add(dA, dAT, 0, 0)                       # pattern only: A' carries no values
    elif order == 2:
        dAT2 = fkeep_device(dAT, KEEP_SHORTCOL, float(dense))
        dA2 = cs_transpose(dAT2, False)
        dC = mult(dAT2, dA2)
    else:
        dC = mult(dAT, dA)
    dC = fkeep_device(dC, KEEP_OFFDIAG)
    if tmp:
        dA.free()
    return dC.download(trim=True) if host else dC


def _mult_ordered(dA, dB):
    out = C.c_void_p()
    _lib.check(_lib.lib().csb200_multiply_ordered(dA._h, dB._h, C.byref(out)), "cs_multiply")
    return out.value


def cs_pinv(p, n):
    """pinv[p[k]] = k (csparse.py:1696-1710); None if p is None.  Host helper."""
    if p is None:
        return None
    pinv = [0] * n
    for k in range(n):
        pinv[p[k]] = k
    return pinv


def cs_permute(A, pinv, q, values):
    """C = P A Q (csparse.py:1666-1693): column k of C is column q[k] of A, row i of A is row
    pinv[i] of C; pinv / q may be None.  None unless compressed-column."""
    if not CS_CSC(A):
        return None
    dA, tmp = _as_device(A)
    try:
        if pinv is not None:
            _need(pinv, dA.m, "cs_permute: pinv")
        if q is not None:
            _need(q, dA.n, "cs_permute: q")
    except IndexError:
        if tmp:
            dA.free()
        raise
    pv = None if pinv is None else _i32(pinv, dA.m)
    qv = None if q is None else _i32(q, dA.n)
    out = C.c_void_p()
    nzmax = max(dA.nnz, 1)                      # cs_spalloc(m, n, Ap[n], ...) (csparse.py:1681)
    try:
        _lib.check(_lib.lib().csb200_permute(dA._h, _ptr(pv), _ptr(qv), 1 if values else 0, C.byref(out)), "cs_permute")
    finally:
        if tmp:
            dA.free()
    dC = DeviceMatrix(out.value)
    if not tmp:
        return dC
    return _padded(dC, nzmax, _numpy_backed(A))


def cs_symperm(A, pinv, values):
    """C = P A P' for a symmetric A whose upper triangular part is stored/used
    (csparse.py:2220-2255).  C.nzmax = max(nnz(A), 1) as the reference allocates it."""
    if not CS_CSC(A):
        return None
    nzmax = max(_nnz_of(A), 1)
    dA, tmp = _as_device(A)
    if pinv is not None and len(pinv) < dA.n:
        if tmp:
            dA.free()
        raise IndexError(f"cs_symperm: pinv: {len(pinv)} entries, {dA.n} needed")
    pv = None if pinv is None else _i32(pinv, dA.n)
    out = C.c_void_p()
    _lib.check(_lib.lib().csb200_symperm(dA._h, _ptr(pv), 1 if values else 0, C.byref(out)), "cs_symperm")
    dC = DeviceMatrix(out.value)
    if not tmp:
        return dC
    dA.free()
    return _padded(dC, nzmax, _numpy_backed(A))


def set_stream(cuda_stream: int = 0):
    """CUDA stream (raw handle, e.g. torch.cuda.current_stream().cuda_stream) used by this
    thread's subsequent calls; 0 selects the legacy default stream."""
    _lib.check(_lib.lib().csb200_set_stream(C.c_void_p(cuda_stream or None)))


def set_device(device: int):
    _lib.check(_lib.lib().csb200_set_device(int(device)), "set_device")


def synchronize():
    _lib.check(_lib.lib().csb200_synchronize(), "synchronize")


def gaxpy_host(m, n, Ap, Ai, Ax, x, y):
    """One-shot cs_gaxpy on host buffers given as raw addresses or numpy arrays
    (csb200_gaxpy_host: upload + CSR build + SpMV + download of y)."""
    def ad(a):
        return C.c_void_p(a) if isinstance(a, int) else _ptr(a)
    _lib.check(_lib.lib().csb200_gaxpy_host(m, n, ad(Ap), ad(Ai), ad(Ax), ad(x), ad(y)), "gaxpy_host")


def force_multiply_path(path: Optional[str]):
    """None / "auto": device-matrix products may use the blocked numeric kernel (rows of a column
    block by block); "ordered": always the reference's discovery order; "blocked_v1" / "blocked_v2":
    automatic with that version of the blocked numeric kernel (A/B measurements); "no_templates":
    automatic without the pattern-class templates; "templates": templates tried at any size;
    "templates_percol": the same without the lane-per-column kernel (one warp per column only)."""
    code = {None: 0, "auto": 0, "ordered": 1, "blocked_v1": 2, "blocked_v2": 3, "no_templates": 4,
            "templates": 5, "templates_percol": 6}[path]
    _lib.check(_lib.lib().csb200_multiply_force_path(code))


def last_multiply_templated() -> int:
    """Columns the last cs_multiply on this thread formed from pattern-class templates."""
    return int(_lib.lib().csb200_multiply_last_templated())


def last_multiply_flops() -> int:
    """Multiply-adds performed by the last cs_multiply on this thread."""
    return int(_lib.lib().csb200_multiply_last_flops())


def launch_count() -> int:
    """Kernels launched by libcsparse_b200.so since it was loaded."""
    return int(_lib.lib().csb200_launch_count())
