#!/usr/bin/env python
"""Development probe (not the contract bench): device-resident timings of the four
kernels on the BASELINE configs, CUDA events, printed as JSON lines."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csparse_cuda as cc
from csparse_cuda import synth

PEAK = 6456.5


ONCE = False


def timeit(fn, warm=3, iters=10):
    if ONCE:
        warm, iters = 1, 1
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def report(name, nbytes, med, best, **kw):
    gbs = nbytes / (med * 1e-3) / 1e9
    print(json.dumps({"what": name, "ms_median": round(med, 4), "ms_best": round(best, 4),
                      "GBs": round(gbs, 1), "frac_of_measured_peak": round(gbs / PEAK, 3), **kw}), flush=True)


ONLY = None
MUL_PATHS = (None,)


def run_matrix(tag, m, n, p, i, x, do_mul, plans=(None,)):
    nnz = len(i)
    if ONLY == "transpose":
        do_mul, plans = False, ()
    if ONLY in ("multiply", "tm"):
        plans = ()
    t0 = time.time()
    dA = cc.from_arrays(m, n, p, i, x)
    print(json.dumps({"what": tag + " upload", "s": round(time.time() - t0, 2), "nnz": nnz}), flush=True)
    holder = {}

    def tr():
        holder["c"] = cc.cs_transpose(dA, True)
    if ONLY != "multiply":
        for path in (None, "bucket", "bucket_fused"):
            cc.force_transpose_path(path)
            try:
                med, best = timeit(tr, 2, 7)
                took = cc.last_transpose_path()
            finally:
                cc.force_transpose_path(None)
            report(f"{tag} cs_transpose[{took}{' fused' if path == 'bucket_fused' else ''}]", synth.transpose_bytes(m, n, nnz), med, best)
            if took == "radix":
                break
    holder.clear()
    xv = torch.randn(n, dtype=torch.float64, device="cuda")
    yv = torch.randn(m, dtype=torch.float64, device="cuda")
    for plan in plans:
        dA.force_gaxpy_plan(plan)
        dA.prepare_gaxpy()
        med, best = timeit(lambda: dA.gaxpy_dev(xv.data_ptr(), yv.data_ptr()), 3, 20)
        report(f"{tag} cs_gaxpy[{dA.gaxpy_plan()}]", synth.gaxpy_bytes(m, n, nnz), med, best,
               gflops=round(2 * nnz / (med * 1e-3) / 1e9, 1))
    if do_mul:
        def mul():
            holder.clear()                   # one 3 GB result alive at a time (keeps the memory pool steady)
            holder["c"] = cc.cs_multiply(dA, dA)
        for path in MUL_PATHS:
            cc.force_multiply_path(path)
            try:
                med, best = timeit(mul, 1, 5)
            finally:
                cc.force_multiply_path(None)
            C = holder["c"]
            report(f"{tag} cs_multiply A*A [{path or 'auto'}]", synth.multiply_bytes(nnz, nnz, C.nnz, n, n), med, best,
                   nnzC=C.nnz, nnzC_per_s=round(C.nnz / (med * 1e-3), 1), madds=cc.last_multiply_flops())
            holder.clear()
    dA.free()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--lap", type=int, default=4096)
    ap.add_argument("--st", type=int, default=128)
    ap.add_argument("--rmat", type=int, default=20)
    ap.add_argument("--only", default=None, choices=[None, "transpose", "multiply", "tm"])
    ap.add_argument("--once", action="store_true", help="one warm-up + one timed call per op (for ncu launch lists)")
    ap.add_argument("--mul-paths", default="auto", help="comma list of auto,blocked_v1,blocked_v2,ordered")
    a = ap.parse_args()
    ONCE = a.once
    ONLY = a.only
    MUL_PATHS = tuple(None if t == 'auto' else t for t in a.mul_paths.split(','))
    torch.cuda.init()
    cc.set_stream(torch.cuda.current_stream().cuda_stream)
    n = 1 << 24
    c = torch.randint(0, 9, (n,), dtype=torch.int32, device="cuda")
    pp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    import ctypes as C
    from csparse_cuda import _lib
    tot = C.c_int64()
    med, best = timeit(lambda: _lib.lib().csb200_cumsum_dev(C.c_void_p(pp.data_ptr()), C.c_void_p(c.data_ptr()), n, C.byref(tot)), 2, 10)
    report("cs_cumsum n=2^24 (incl. D2H of total)", synth.cumsum_bytes(n), med, best)
    if a.lap:
        run_matrix(f"lap2d {a.lap}", *synth.lap2d(a.lap), do_mul=a.lap <= 2048, plans=("stream", "stream_ld", "merge"))
    if a.st:
        run_matrix(f"st27 {a.st}", *synth.st27(a.st), do_mul=True, plans=("stream", "stream_ld"))
    if a.rmat:
        run_matrix(f"rmat {a.rmat}", *synth.rmat(a.rmat, 16), do_mul=False, plans=("merge", "stream"))
