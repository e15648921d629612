#!/usr/bin/env python
"""Development probe: cs_transpose (radix path) and cs_gaxpy (merge path) on a GPU-generated
R-MAT matrix; median of CUDA-event timings, one result alive at a time."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csparse_cuda as cc
from csparse_cuda import synth

ap = argparse.ArgumentParser()
ap.add_argument("--scale", type=int, default=24)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--no-gaxpy", action="store_true")
ap.add_argument("--no-transpose", action="store_true")
ap.add_argument("--plans", default="auto", help="comma list of auto,merge,split,stream")
a = ap.parse_args()
torch.cuda.init()
cc.set_stream(torch.cuda.current_stream().cuda_stream)
m, n, tp, ti, tx = synth.rmat_torch(a.scale, 16)
nnz = int(ti.numel())
dA = cc.from_device(m, n, tp.data_ptr(), ti.data_ptr(), tx.data_ptr())
torch.cuda.synchronize()
del tp, ti, tx
torch.cuda.empty_cache()


def timed(fn, warm, iters):
    ts = []
    for k in range(warm + iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if k >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def tr():
    c = cc.cs_transpose(dA, True)
    c.free()


if not a.no_transpose:
    med, best = timed(tr, 2, a.iters)
    b = synth.transpose_bytes(m, n, nnz)
    print(json.dumps({"what": f"rmat {a.scale} cs_transpose", "nnz": nnz, "ms_median": med, "ms_best": best,
                      "GBs": b / med / 1e6}), flush=True)
if not a.no_gaxpy:
    xv = torch.randn(n, dtype=torch.float64, device="cuda")
    yv = torch.randn(m, dtype=torch.float64, device="cuda")
    for plan in a.plans.split(","):
        dA.force_gaxpy_plan(None if plan == "auto" else plan)
        dA.prepare_gaxpy()
        med, best = timed(lambda: dA.gaxpy_dev(xv.data_ptr(), yv.data_ptr()), 3, 4 * a.iters)
        b = synth.gaxpy_bytes(m, n, nnz)
        print(json.dumps({"what": f"rmat {a.scale} cs_gaxpy[{dA.gaxpy_plan()}]", "ms_median": med, "ms_best": best,
                          "GBs": b / med / 1e6, "frac_of_measured_peak": b / med / 1e6 / 6456.5}), flush=True)
