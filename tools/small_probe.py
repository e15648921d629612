#!/usr/bin/env python
"""Development probe: BASELINE configs 1-2 (bcsstk01 / bcsstk16 from tests/golden): latency of
cs_transpose, cs_multiply A*A', cs_gaxpy on device handles and through host `cs` objects."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import csparse_cuda as cc
from oracle import oracle as orc

torch.cuda.init()
cc.set_stream(torch.cuda.current_stream().cuda_stream)

def timeit(fn, warm=3, iters=20):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))

for name in ("bcsstk01", "bcsstk16"):
    z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".npz"))
    meta = json.loads(str(z["meta"]))
    m, n = meta["A"]["m"], meta["A"]["n"]
    p, i, x = z["A_p"], z["A_i"], z["A_x"]
    dA = cc.from_arrays(m, n, p, i, x)
    dAT = cc.cs_transpose(dA, True)
    h = {}
    def tr(): h["t"] = cc.cs_transpose(dA, True)
    def mul(): h["c"] = cc.cs_multiply(dA, dAT)
    xv = torch.randn(n, dtype=torch.float64, device="cuda"); yv = torch.zeros(m, dtype=torch.float64, device="cuda")
    dA.prepare_gaxpy()
    def gx(): dA.gaxpy_dev(xv.data_ptr(), yv.data_ptr())
    out = {"matrix": name, "nnz": int(len(i)),
           "transpose_dev_ms": timeit(tr), "multiply_dev_ms": timeit(mul), "gaxpy_dev_ms": timeit(gx)}
    out["nnzC"] = h["c"].nnz
    # host cs objects (numpy-backed): upload + kernels + download every call
    A = cc.cs(); A.m, A.n, A.nzmax, A.nz = m, n, len(i), -1; A.p, A.i, A.x = p.copy(), i.copy(), x.copy()
    def tr_h(): h["th"] = cc.cs_transpose(A, True)
    out["transpose_host_ms"] = timeit(tr_h, 2, 10)
    AT = h["th"]
    def mul_h(): h["ch"] = cc.cs_multiply(A, AT)
    out["multiply_host_ms"] = timeit(mul_h, 2, 10)
    # the C oracle on the same inputs
    oA = orc.csc(m, n, p, i, x)
    t0 = time.perf_counter(); oT = orc.cs_transpose(oA, True); out["transpose_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    t0 = time.perf_counter(); oC = orc.cs_multiply(oA, oT); out["multiply_oracle_ms"] = (time.perf_counter() - t0) * 1e3
    print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in out.items()}), flush=True)
