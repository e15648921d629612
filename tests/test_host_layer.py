"""CPU-only checks of the host layer and the C-ABI boundary: the library loads and
exports every symbol include/csparse_b200.h declares, argument-shape errors map to
the reference's sentinels without touching a GPU, and the product package never
reaches into oracle/."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

import csparse_cuda as cc
from csparse_cuda import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "csparse_b200.h")).read()
    return sorted(set(re.findall(r"CSB200_API\s+[\w\s\*]+?\b(csb200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 25
    L = _lib.lib()
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/csparse_b200.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), "ctypes prototypes and header out of sync"
    assert L.csb200_version() >= 100
    assert isinstance(_lib.last_error(), str)


def test_header_cites_the_reference_for_each_entry_point():
    src = open(os.path.join(ROOT, "include", "csparse_b200.h")).read()
    for ref in ("csparse.py:767-784", "csparse.py:2292-2315", "csparse.py:1199-1213", "csparse.py:1608-1642",
                "csparse.py:37-54"):
        assert ref in src


def test_cs_object_and_predicates():
    A = cc.cs()
    assert (A.nzmax, A.m, A.n, A.p, A.i, A.x, A.nz) == (0, 0, 0, [], [], [], 0)
    assert cc.CS_TRIPLET(A) and not cc.CS_CSC(A)
    A.nz = -1
    assert cc.CS_CSC(A) and not cc.CS_TRIPLET(A)
    assert not cc.CS_CSC(None) and not cc.CS_TRIPLET(None)


def test_sentinels_need_no_gpu():
    """Argument-shape errors return None / False / -1 like the reference
    (csparse.py:777, :1207-1208, :1616-1619, :2299-2300) before any device work."""
    T = cc.cs(); T.nz = 2; T.m = T.n = 2
    assert cc.cs_transpose(T, True) is None and cc.cs_transpose(None, False) is None
    assert cc.cs_multiply(T, T) is None and cc.cs_multiply(None, None) is None
    assert cc.cs_gaxpy(T, [1.0], [1.0]) is False and cc.cs_gaxpy(None, [1.0], [1.0]) is False
    A = cc.cs(); A.nz = -1; A.m, A.n = 2, 3; A.p, A.i, A.x = [0, 0, 0, 0], [0], [0.0]
    assert cc.cs_gaxpy(A, None, [0.0, 0.0]) is False and cc.cs_gaxpy(A, [0.0] * 3, None) is False
    assert cc.cs_multiply(A, A) is None                       # A.n != B.m
    assert cc.cs_cumsum(None, [1], 1) == -1 and cc.cs_cumsum([0, 0], None, 1) == -1
    P = cc.cs(); P.nz = -1; P.m = P.n = 1; P.p, P.i, P.x = [0, 1], [0], None
    with pytest.raises(TypeError):
        cc.cs_gaxpy(P, [1.0], [1.0])


def test_next_row_sentinels_need_no_gpu():
    """cs_add :173-176, cs_norm :1655, cs_compress :654, cs_dupl :1044, cs_fkeep :1182,
    cs_permute :1680, cs_symperm :2229, cs_amd :229 -- argument-shape errors before any device work"""
    T = cc.cs(); T.nz = 2; T.m = T.n = 2
    A = cc.cs(); A.nz = -1; A.m, A.n = 2, 3; A.p, A.i, A.x = [0, 0, 0, 0], [0], [0.0]
    B = cc.cs(); B.nz = -1; B.m, B.n = 3, 3; B.p, B.i, B.x = [0, 0, 0, 0], [0], [0.0]
    assert cc.cs_add(T, A, 1, 1) is None and cc.cs_add(A, None, 1, 1) is None and cc.cs_add(A, B, 1, 1) is None
    assert cc.cs_norm(None) == -1 and cc.cs_norm(T) == -1
    P = cc.cs(); P.nz = -1; P.m = P.n = 1; P.p, P.i, P.x = [0, 1], [0], None
    assert cc.cs_norm(P) == -1
    assert cc.cs_compress(A) is None and cc.cs_compress(None) is None
    assert cc.cs_dupl(T) is False and cc.cs_dupl(None) is False
    assert cc.cs_fkeep(T, cc.KEEP_NONZERO, None) == -1 and cc.cs_dropzeros(None) == -1 and cc.cs_droptol(T, 1.0) == -1
    assert cc.cs_permute(T, None, None, True) is None and cc.cs_symperm(None, None, True) is None
    assert cc.cs_amd_matrix(0, A) is None and cc.cs_amd_matrix(4, A) is None and cc.cs_amd_matrix(1, T) is None
    assert cc.cs_pinv(None, 3) is None and cc.cs_pinv([2, 0, 1], 3) == [1, 2, 0]
    with pytest.raises(TypeError):
        cc.cs_dupl(P)
    with pytest.raises(NotImplementedError):
        cc.cs_fkeep(A, cc.cs_ifkeep(), None)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    A = cc.cs(); A.nz = -1; A.m = A.n = 2; A.nzmax = 2
    A.p, A.i, A.x = [0, 1, 2], [0, 1], [1.0, 2.0]
    for call in (lambda: cc.cs_transpose(A, True), lambda: cc.cs_multiply(A, A),
                 lambda: cc.cs_gaxpy(A, [1.0, 1.0], [0.0, 0.0]), lambda: cc.cs_cumsum([0, 0, 0], [1, 2], 2)):
        with pytest.raises(cc.CSparseCudaError):
            call()


def test_product_never_imports_the_oracle():
    for path in glob.glob(os.path.join(ROOT, "csparse_cuda", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".h")):
            text = open(path, errors="replace").read()
            assert "oracle" not in text.lower(), f"{path} mentions the oracle"


def test_marshalling():
    assert cc._i32([1, 2, 3, 4], 3).tolist() == [1, 2, 3] and cc._i32([1, 2, 3], 3).dtype == np.int32
    a = np.arange(5, dtype=np.int32)
    assert cc._i32(a, 5) is a or np.shares_memory(cc._i32(a, 5), a)          # zero-copy for int32 numpy
    assert cc._f64([1, 2.5], 2).tolist() == [1.0, 2.5]                       # ints are promoted
    import array
    assert cc._i32(array.array("i", [4, 5, 6]), 2).tolist() == [4, 5]
    with pytest.raises(OverflowError):
        cc._i32([2 ** 31], 1)


def test_synthetic_generators():
    m, n, p, i, x = synth.lap2d(16)
    assert m == n == 256 and p[-1] == len(i) == 5 * 256 - 4 * 16
    assert np.all(np.diff(i.astype(np.int64))[np.diff(np.repeat(np.arange(n), np.diff(p))) == 0] > 0)
    d = np.zeros((m, n)); d[i, np.repeat(np.arange(n), np.diff(p))] = x
    assert np.array_equal(d, d.T) and np.all(np.diag(d) == 4.0) and np.all(d.sum(1) >= 0)
    # column slabs of a rectangular grid tile the whole matrix
    full = synth.lap2d_cols(8, 24, 0, 192)
    parts = [synth.lap2d_cols(8, 24, a, b) for a, b in ((0, 64), (64, 128), (128, 192))]
    assert np.array_equal(np.concatenate([q[3] for q in parts]), full[3])
    assert sum(len(q[3]) for q in parts) == 5 * 192 - 2 * 8 - 2 * 24
    m, n, p, i, x = synth.st27(5)
    assert p[-1] == len(i) == (3 * 5 - 2) ** 3 and x.min() >= 0.5 and x.max() <= 1.5
    m, n, p, i, x = synth.rmat(8, 4)
    cols = np.repeat(np.arange(n), np.diff(p))
    key = cols.astype(np.int64) * m + i
    assert np.all(np.diff(key) > 0)                      # canonical: sorted, no duplicates
    assert synth.gaxpy_bytes(16777216, 16777216, 83869696) == 1476198404
    assert synth.transpose_bytes(16777216, 16777216, 83869696) == 2147090440
    assert synth.multiply_bytes(55742968, 55742968, 254840104, 2097152, 2097152) == 4421078316


def test_path_controls_need_no_gpu():
    """the test / A-B switches of the C ABI only set library state: documented names map to the
    documented codes, anything else is refused, and nothing touches a device"""
    from csparse_cuda import _lib
    L = _lib.lib()
    for name in (None, "auto", "radix", "bucket"):
        cc.force_transpose_path(name)
    cc.force_transpose_path(None)
    assert cc.last_transpose_path() == "trivial"                  # no transpose ran on this thread
    with pytest.raises(KeyError):
        cc.force_transpose_path("fastest")
    assert L.csb200_transpose_force_path(7) != 0 and L.csb200_transpose_force_path(-1) != 0
    for name in (None, "auto", "ordered", "blocked_v1", "blocked_v2"):
        cc.force_multiply_path(name)
    cc.force_multiply_path(None)
    with pytest.raises(KeyError):
        cc.force_multiply_path("blocked_v9")
    assert L.csb200_multiply_force_path(9) != 0
    assert L.csb200_transpose_force_path(0) == 0 and L.csb200_multiply_force_path(0) == 0


def test_numpy_backed_detection():
    """Results keep the container kind of the operands: numpy-backed cs objects are recognised by their
    row-index array; list-backed ones (the reference's own kind) and device handles are not."""
    import numpy as np
    import csparse_cuda as cc
    A = cc.cs()
    A.m = A.n = 2
    A.nzmax, A.nz = 2, -1
    A.p, A.i, A.x = [0, 1, 2], [0, 1], [1.0, 2.0]
    B = cc.cs()
    B.m = B.n = 2
    B.nzmax, B.nz = 2, -1
    B.p, B.i, B.x = np.array([0, 1, 2], np.int32), np.array([0, 1], np.int32), np.array([1.0, 2.0])
    assert cc._numpy_backed(A) is False
    assert cc._numpy_backed(B) is True
    assert cc._numpy_backed(A, B) is True
