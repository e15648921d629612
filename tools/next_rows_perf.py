#!/usr/bin/env python
"""Development probe: device-resident timings of the rows either side of the hot path
(cs_compress, cs_add, cs_fkeep, cs_permute, cs_symperm, cs_norm, cs_dupl) on lap2d k x k."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import csparse_cuda as cc
from csparse_cuda import synth, _lib

PEAK = 6456.5
ap = argparse.ArgumentParser()
ap.add_argument("--k", type=int, default=4096)
a = ap.parse_args()
torch.cuda.init()
cc.set_stream(torch.cuda.current_stream().cuda_stream)
m, n, p, i, x = synth.lap2d(a.k)
nnz = len(i)
dA = cc.from_arrays(m, n, p, i, x)


def timed(fn, warm=2, iters=5):
    ts = []
    for k in range(warm + iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        if hasattr(r, "free"):
            r.free()
        if k >= warm:
            ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def report(name, nbytes, ms, **kw):
    print(json.dumps({"what": f"lap2d {a.k} {name}", "ms": round(ms, 4), "GBs": round(nbytes / ms / 1e6, 1),
                      "frac_of_measured_peak": round(nbytes / ms / 1e6 / PEAK, 3), **kw}), flush=True)


# triplets on the device, shuffled (cs_compress does not assume any order)
cols = torch.repeat_interleave(torch.arange(n, dtype=torch.int32, device="cuda"),
                               torch.from_numpy(np.diff(p).astype(np.int64)).cuda())
rows = torch.from_numpy(i).cuda()
vals = torch.from_numpy(x).cuda()
perm = torch.randperm(nnz, device="cuda")
tj, ti, tx = cols[perm].contiguous(), rows[perm].contiguous(), vals[perm].contiguous()
del cols, rows, vals, perm


def compress():
    out = C.c_void_p()
    _lib.check(_lib.lib().csb200_compress_dev(m, n, nnz, C.c_void_p(ti.data_ptr()), C.c_void_p(tj.data_ptr()),
                                              C.c_void_p(tx.data_ptr()), C.byref(out)))
    return cc.DeviceMatrix(out.value)


report("cs_compress (shuffled triplets)", 32 * nnz + 4 * (n + 1), timed(compress), nnz=nnz)
del ti, tj, tx
dAT = cc.cs_transpose(dA, True)
report("cs_add A+A'", 12 * (2 * nnz + nnz) + 12 * (n + 1), timed(lambda: cc.cs_add(dA, dAT, 1.0, 1.0)))
report("cs_norm", 8 * nnz + 4 * (n + 1), timed(lambda: cc.cs_norm(dA)))
report("cs_fkeep offdiag", 12 * nnz + 12 * (nnz - n) + 8 * (n + 1), timed(lambda: cc.fkeep_device(dA, cc.KEEP_OFFDIAG)))
pinv = np.random.default_rng(0).permutation(n).astype(np.int32)
q = np.random.default_rng(1).permutation(n).astype(np.int32)
report("cs_permute (incl. H2D of pinv, q)", 24 * nnz + 16 * (n + 1), timed(lambda: cc.cs_permute(dA, pinv, q, True)))
report("cs_symperm (incl. H2D of pinv)", 12 * nnz + 12 * ((nnz + n) // 2), timed(lambda: cc.cs_symperm(dA, pinv, True)))
report("cs_dupl", 24 * nnz + 8 * (n + 1), timed(lambda: cc.dupl_device(dA)))
