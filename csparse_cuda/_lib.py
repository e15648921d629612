"""ctypes binding of libcsparse_b200.so (include/csparse_b200.h).

The library is loaded on first use.  There is no CPU fallback: if the shared
library is missing, or no CUDA device is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CSPARSE_B200_LIB", os.path.join(_HERE, "libcsparse_b200.so"))

OK, ERR_ARG, ERR_CUDA, ERR_OVERFLOW, ERR_INDEX, ERR_NOMEM = range(6)

i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
i64p = C.POINTER(C.c_int64)
intp = C.POINTER(C.c_int)
mat_t = C.c_void_p
matp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); mirrors include/csparse_b200.h one to one
PROTOTYPES = {
    "csb200_version": (C.c_int, []),
    "csb200_last_error": (C.c_char_p, []),
    "csb200_device_count": (C.c_int, [intp]),
    "csb200_set_device": (C.c_int, [C.c_int]),
    "csb200_set_stream": (C.c_int, [C.c_void_p]),
    "csb200_synchronize": (C.c_int, []),
    "csb200_sm_count": (C.c_int, [intp]),
    "csb200_launch_count": (C.c_int64, []),
    "csb200_cumsum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, i64p]),
    "csb200_cumsum_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, i64p]),
    "csb200_mat_upload": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, matp]),
    "csb200_mat_from_dev": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, matp]),
    "csb200_mat_dims": (C.c_int, [mat_t, i32p, i32p, i64p, intp]),
    "csb200_mat_download": (C.c_int, [mat_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "csb200_mat_dev_ptrs": (C.c_int, [mat_t, matp, matp, matp]),
    "csb200_mat_col_slice": (C.c_int, [mat_t, C.c_int32, C.c_int32, matp]),
    "csb200_mat_free": (C.c_int, [mat_t]),
    "csb200_transpose": (C.c_int, [mat_t, C.c_int, matp]),
    "csb200_transpose_force_path": (C.c_int, [C.c_int]),
    "csb200_transpose_last_path": (C.c_int, []),
    "csb200_transpose_host": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]),
    "csb200_gaxpy": (C.c_int, [mat_t, C.c_void_p, C.c_void_p]),
    "csb200_gaxpy_dev": (C.c_int, [mat_t, C.c_void_p, C.c_void_p]),
    "csb200_gaxpy_t_dev": (C.c_int, [mat_t, C.c_void_p, C.c_void_p]),
    "csb200_gaxpy_host": (C.c_int, [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "csb200_gaxpy_prepare": (C.c_int, [mat_t]),
    "csb200_gaxpy_plan": (C.c_int, [mat_t, intp]),
    "csb200_gaxpy_force_plan": (C.c_int, [mat_t, C.c_int]),
    "csb200_halo_create": (C.c_int, [C.c_int64, matp]),
    "csb200_halo_window": (C.c_int, [C.c_void_p, matp]),
    "csb200_halo_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "csb200_halo_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "csb200_halo_connect_local": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64]),
    "csb200_gaxpy_halo_dev": (C.c_int, [mat_t, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "csb200_gaxpy_halo": (C.c_int, [mat_t, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p,
                                    C.c_void_p]),
    "csb200_halo_status": (C.c_int, [C.c_void_p, intp]),
    "csb200_halo_free": (C.c_int, [C.c_void_p]),
    "csb200_multiply": (C.c_int, [mat_t, mat_t, matp]),
    "csb200_multiply_ordered": (C.c_int, [mat_t, mat_t, matp]),
    "csb200_multiply_force_path": (C.c_int, [C.c_int]),
    "csb200_multiply_last_flops": (C.c_int64, []),
    "csb200_multiply_last_templated": (C.c_int64, []),
    "csb200_add": (C.c_int, [mat_t, mat_t, C.c_double, C.c_double, matp]),
    "csb200_add_force_path": (C.c_int, [C.c_int]),
    "csb200_norm": (C.c_int, [mat_t, f64p]),
    "csb200_compress": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, matp]),
    "csb200_compress_dev": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, matp]),
    "csb200_dupl": (C.c_int, [mat_t, matp]),
    "csb200_fkeep": (C.c_int, [mat_t, C.c_int, C.c_double, matp]),
    "csb200_permute": (C.c_int, [mat_t, C.c_void_p, C.c_void_p, C.c_int, matp]),
    "csb200_symperm": (C.c_int, [mat_t, C.c_void_p, C.c_int, matp]),
}

_lib = None


class CSparseCudaError(RuntimeError):
    """CUDA / driver failure inside libcsparse_b200.so (no CPU fallback exists)."""


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CSparseCudaError(
                f"{LIB_PATH} not found: build it with `python -m csparse_cuda.build` "
                "(nvcc, sm_100a). csparse_cuda has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    msg = lib().csb200_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(status: int, what: str = "", allow_arg: bool = False) -> int:
    """Map a csb200_status to the Python exception the host layer documents.
    ERR_ARG raises ValueError unless ``allow_arg`` (the few callers that turn it into the
    reference's sentinel or into their own exception ask for it back)."""
    if status == OK or (status == ERR_ARG and allow_arg):
        return status
    msg = f"{what}: {last_error()}" if what else last_error()
    if status == ERR_ARG:
        raise ValueError(msg)
    if status == ERR_INDEX:
        raise ValueError(msg)
    if status == ERR_OVERFLOW:
        raise OverflowError(msg)
    if status == ERR_NOMEM:
        raise MemoryError(msg)
    raise CSparseCudaError(msg)
