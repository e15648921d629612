#!/usr/bin/env python
"""Diagnostic: per-call time of cs_transpose (bucket path) on lap2d 4096^2 in back-to-back batches
versus one synchronised call at a time, before and after the mirror path has run in the process."""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csparse_cuda as cc
from csparse_cuda import synth

torch.cuda.init()
cc.set_stream(torch.cuda.current_stream().cuda_stream)
m, n, p, i, x = synth.lap2d(4096)
dA = cc.from_arrays(m, n, p, i, x)
hold = {}
ev = lambda: torch.cuda.Event(enable_timing=True)

def batches(path, clear_first, nb=3, iters=5):
    cc.force_transpose_path(path)
    out = []
    try:
        for _ in range(nb):
            e0, e1 = ev(), ev(); e0.record()
            for _ in range(iters):
                if clear_first: hold.clear()
                hold["c"] = cc.cs_transpose(dA, True)
            e1.record(); torch.cuda.synchronize()
            out.append(round(e0.elapsed_time(e1) / iters, 3))
    finally:
        cc.force_transpose_path(None)
    return out

def percall(path, iters=5):
    cc.force_transpose_path(path)
    out = []
    try:
        for _ in range(iters):
            e0, e1 = ev(), ev(); e0.record()
            hold["c"] = cc.cs_transpose(dA, True)
            e1.record(); torch.cuda.synchronize()
            out.append(round(e0.elapsed_time(e1), 3))
    finally:
        cc.force_transpose_path(None)
    return out

print("bucket back-to-back (fresh)      ", batches("bucket", False), flush=True)
print("bucket per call                  ", percall("bucket"), flush=True)
print("mirror back-to-back              ", batches(None, False), flush=True)
print("bucket back-to-back after mirror ", batches("bucket", False), flush=True)
print("bucket per call after mirror     ", percall("bucket"), flush=True)
print("bucket back-to-back, clear first ", batches("bucket", True), flush=True)
hold.clear()
print("mem", torch.cuda.mem_get_info(), flush=True)
