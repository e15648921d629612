#!/usr/bin/env python
"""bench.py -- the measurement contract for the csparse_cuda hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl reference]

One "step" is one pass of the hot path over one resident input.  The default
workload is BASELINE.json's headline: cs_gaxpy (y += A*x) on the 2-D 5-point
Laplacian 4096^2 (n = 16 777 216, nnz = 83 869 696), metric "cs_gaxpy HBM GB/s" =
algorithmic bytes (12 nnz + 4(n+1) + 8n + 16m) / time.  Other
workloads: multiply_st27 (cs_multiply A*A, 27-point stencil 128^3, nnz(C)/s),
transpose_lap2d, gaxpy_rmat (R-MAT 2^scale rows, merge-path kernel; N > 1: row blocks balanced by
nonzeros, x all-gathered every step).

N > 1 (launched by torchrun, one rank per GPU, NCCL): weak scaling -- every rank
owns a 4096 x 4096 slab (16.7 M rows) of a 4096 x 4096N grid; per step each rank
exchanges its x halo with its neighbours and runs the local row-block SpMV.

`--impl reference` times the reference algorithm on the host CPU (the C oracle
port of csparse.py's loops, single thread -- the reference has no parallelism).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="gaxpy_lap2d",
                    choices=["gaxpy_lap2d", "multiply_st27", "transpose_lap2d", "gaxpy_rmat"])
    ap.add_argument("--k", type=int, default=0, help="grid edge (lap2d: 4096, st27: 128) or R-MAT scale (24)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--halo", default="fused", choices=["fused", "nccl"],
                    help="N > 1 halo exchange of cs_gaxpy: pulled over NVLink inside the one SpMV launch, or batched NCCL send/recv")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads in the N=1 default run")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region, for the GPUs of this job
    (rank 0 samples all of them: eight ranks each starting their own nvidia-smi took longer to
    produce a first line than the timed region lasted)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, indices, enabled: bool = True):
        self.indices = [indices] if isinstance(indices, int) else list(indices)
        self.enabled, self.proc, self.lines, self.mark = enabled, None, [], 0

    def launch(self):
        """Start nvidia-smi early (before the warm-up): its start-up takes a second or more."""
        if not self.enabled or self.proc:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", ",".join(str(i) for i in self.indices), "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def start(self):
        """Mark the beginning of the timed region: only samples taken from here on count."""
        self.launch()
        if self.proc:
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:     # first line = nvidia-smi is up
                time.sleep(0.02)
        self.mark = len(self.lines)

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.enabled:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampled on rank 0 only"], "samples": 0}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        time.sleep(0.15)
        lines = self.lines[self.mark:]
        self.mark = len(self.lines)
        sm, mx, reasons = [], [], set()
        for ln in lines:
            f = [t.strip() for t in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "gpus_sampled": len(self.indices)}

    def close(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
            self.proc = None


# ------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host CPU
# ------------------------------------------------------------------------------------

def cpu_workload(workload: str, k: int):
    """(callable running ONE pass on the CPU, units per pass, description, unit-of-metric)"""
    from csparse_cuda import synth
    from oracle import oracle as orc
    if workload == "gaxpy_lap2d":
        m, n, p, i, x = synth.lap2d(k)
        A = orc.csc(m, n, p, i, x)
        xv, y = synth.vectors(m, n)
        return (lambda: orc.cs_gaxpy(A, xv, y)), synth.gaxpy_bytes(m, n, len(i)) / 1e9, \
            f"full cs_gaxpy pass, lap2d {k}^2 (nnz {len(i)})", "GB/s"
    if workload == "gaxpy_rmat":
        m, n, p, i, x = synth.rmat(k, 16)
        A = orc.csc(m, n, p, i, x)
        xv, y = synth.vectors(m, n)
        return (lambda: orc.cs_gaxpy(A, xv, y)), synth.gaxpy_bytes(m, n, len(i)) / 1e9, \
            f"full cs_gaxpy pass, R-MAT scale {k} (nnz {len(i)})", "GB/s"
    if workload == "transpose_lap2d":
        m, n, p, i, x = synth.lap2d(k)
        A = orc.csc(m, n, p, i, x)
        return (lambda: orc.cs_transpose(A, True)), synth.transpose_bytes(m, n, len(i)) / 1e9, \
            f"full cs_transpose pass, lap2d {k}^2", "GB/s"
    if workload == "multiply_st27":
        m, n, p, i, x = synth.st27(k)
        A = orc.csc(m, n, p, i, x)
        nnzc = (5 * k - 6) ** 3
        return (lambda: orc.cs_multiply(A, A)), float(nnzc), \
            f"full cs_multiply A*A pass, st27 {k}^3 (nnz(C) {nnzc})", "nnz(C)/s"
    raise ValueError(workload)


def time_cpu(fn, units, min_seconds=6.0, min_passes=2, max_passes=200):
    fn()                                   # warm the caches / page in
    t0, passes = time.perf_counter(), 0
    while passes < min_passes or (time.perf_counter() - t0 < min_seconds and passes < max_passes):
        fn()
        passes += 1
    dt = time.perf_counter() - t0
    return units * passes / dt, passes, dt


def default_k(workload: str) -> int:
    return {"gaxpy_lap2d": 4096, "transpose_lap2d": 4096, "multiply_st27": 128, "gaxpy_rmat": 24}[workload]


def cpu_sample_k(workload: str, k: int) -> int:
    """bounded CPU sample of the same family (seconds, not minutes)"""
    if workload == "multiply_st27":
        return min(k, 64)          # 128^3 takes ~10 s/pass and 6 GB in C; 64^3 is the same per-column work
    if workload == "gaxpy_rmat":
        return min(k, 20)
    return k


def run_reference(a):
    """--impl reference: the reference algorithm (oracle C port of csparse.py) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    orc.build()
    k = a.k or default_k(a.workload)
    ks = cpu_sample_k(a.workload, k)
    fn, units, desc, unit = cpu_workload(a.workload, ks)
    for _ in range(max(a.warmup, 1) if ks < 2048 else 1):
        fn()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        fn()
    dt = time.perf_counter() - t0
    value = units * a.steps / dt
    metric = "cs_gaxpy HBM GB/s" if a.workload.startswith("gaxpy") else \
        ("cs_multiply nnz(C)/s" if a.workload == "multiply_st27" else "cs_transpose HBM GB/s")
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True,
        "scaling": a.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": a.workload, "k": k, "cpu_sample_k": ks, "sample": desc},
        "cpu_baseline": {"value": value, "unit": unit, "cores": 1, "kind": "port",
                         "sample": desc + "; oracle/csparse_oracle.c (C restatement of csparse.py loops), 1 thread; "
                                          "the Python reference itself is single-threaded and cannot travel to the GPU box"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


# ------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------

def device_timed(torch, dist, world, fn, steps, warmup):
    """W untimed + exactly K timed steps, barrier + synchronize on both sides, CUDA events,
    max over ranks.  Returns total milliseconds."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; csparse_cuda has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import csparse_cuda as cc
    from csparse_cuda import synth, dist as csd
    cc.set_device(local)
    cc.set_stream(torch.cuda.current_stream().cuda_stream)
    peak, peak_src = peaks()
    k = a.k or default_k(a.workload)
    out = {}
    sampler = ClockSampler(list(range(world)) if world > 1 else local, enabled=(rank == 0))
    sampler.launch()

    if a.workload == "gaxpy_lap2d":
        out = bench_gaxpy_lap2d(a, torch, dist, cc, synth, csd, world, rank, k, peak, peak_src, sampler)
    elif a.workload == "gaxpy_rmat":
        if world == 1:
            out = bench_gaxpy_single(a, torch, cc, synth, "rmat", k, peak, peak_src, sampler)
        else:
            out = bench_gaxpy_rmat_dist(a, torch, dist, cc, synth, csd, world, rank, k, peak, peak_src, sampler)
    elif a.workload == "transpose_lap2d":
        out = bench_transpose(a, torch, cc, synth, k, peak, peak_src, sampler)
    elif a.workload == "multiply_st27":
        out = bench_multiply(a, torch, dist, cc, synth, csd, world, rank, k, peak, peak_src, sampler)

    if world > 1 and a.workload == "gaxpy_lap2d" and not a.no_extra:
        # the other sharded configurations of BASELINE.json on the same N GPUs, every rank taking part
        out["extra"] = extras_dist(a, torch, dist, cc, synth, csd, world, rank, peak, peak_src, out)
    if rank == 0:
        if world == 1 and a.workload == "gaxpy_lap2d" and not a.no_extra:
            out["extra"] = extras(a, torch, cc, synth, peak)
        if world == 1 and not a.no_cpu:
            from oracle import oracle as orc
            orc.build()
            ks = cpu_sample_k(a.workload, k)
            fn, units, desc, unit = cpu_workload(a.workload, ks)
            v, passes, dt = time_cpu(fn, units)
            out["cpu_baseline"] = {"value": v, "unit": unit, "cores": 1, "kind": "port",
                                   "sample": f"{passes} x {desc} in {dt:.1f} s; oracle/csparse_oracle.c, 1 thread "
                                             f"(host has {os.cpu_count()} logical CPUs; the reference is single-threaded)"}
        print(json.dumps(out), flush=True)
    sampler.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def extras_dist(a, torch, dist, cc, synth, csd, world, rank, peak, peak_src, headline):
    """N > 1: driver-visible lines for the sharded configurations the default workload does not
    cover -- BASELINE.json configs 3 (strong-scaled: the fixed 4096^2 matrix split over N GPUs), 4
    (cs_multiply st27 128^3 by column blocks of B; C left distributed, and gathered to rank 0) and 5
    (cs_gaxpy R-MAT 2^24, x all-gathered).  Every rank runs them; a failure on any rank is reported
    instead of sinking the headline, and a watchdog ends a rank that waits for a peer that gave up."""
    import copy
    ex = {}
    deadline = threading.Timer(420.0, lambda: (print(json.dumps(dict(headline, extra=dict(ex, error="extras timed out"))),
                                                   flush=True) if rank == 0 else None, os._exit(0)))
    deadline.daemon = True
    deadline.start()
    quiet = ClockSampler(0, enabled=False)

    def keep(name, r, keys):
        if rank == 0:
            ex[name] = {k: r.get(k) for k in keys if k in r}

    try:
        b = copy.copy(a)
        b.scaling, b.steps, b.warmup = "strong", min(a.steps, 200), max(a.warmup, 3)
        r = bench_gaxpy_lap2d(b, torch, dist, cc, synth, csd, world, rank, 4096, peak, peak_src, quiet)
        keep("cs_gaxpy lap2d 4096^2, strong-scaled over N GPUs", r,
             ("value", "unit", "ms_per_step", "scaling", "n_gpus", "steps", "config", "roofline", "e2e"))
        torch.cuda.empty_cache()
        b = copy.copy(a)
        b.steps, b.warmup = min(a.steps, 50), max(a.warmup, 3)
        r = bench_gaxpy_rmat_dist(b, torch, dist, cc, synth, csd, world, rank, 24, peak, peak_src, quiet)
        keep("cs_gaxpy rmat 2^24, row blocks balanced by nonzeros, x all-gathered", r,
             ("value", "unit", "ms_per_step", "scaling", "n_gpus", "steps", "config", "roofline", "e2e"))
        torch.cuda.empty_cache()
        b = copy.copy(a)
        b.steps, b.warmup = 10, 3
        r = bench_multiply(b, torch, dist, cc, synth, csd, world, rank, 128, peak, peak_src, quiet)
        keep("cs_multiply st27 128^3 A*A, column blocks of B", r,
             ("value", "unit", "ms_per_step", "scaling", "n_gpus", "steps", "config", "roofline"))
    except Exception as e:  # secondary numbers never sink the headline
        ex["error"] = f"rank {rank}: {e!r}"
    deadline.cancel()
    return ex


def bind_numa(local: int):
    """Pin this process to the CPUs next to GPU `local` (sysfs local_cpulist of its PCI function), so
    that pinned host buffers allocated from now on come from that NUMA node.  Returns the node or None."""
    try:
        bdf = subprocess.check_output(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                                      text=True, timeout=20).strip().lower()
        dom, rest = bdf.split(":", 1)
        dev = f"/sys/bus/pci/devices/{dom[-4:]}:{rest}"
        cpus = set()
        with open(dev + "/local_cpulist") as f:
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        if cpus:
            os.sched_setaffinity(0, cpus)
        with open(dev + "/numa_node") as f:
            return int(f.read().strip())
    except Exception:
        return None


def measured_traffic(key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from profiles/traffic.json, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[key]["dram_bytes"]
    except Exception:
        return None


def pinned(torch, arr):
    t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
    return t


def e2e_transpose(torch, cc, m, n, p, i, x, steps=3):
    """cs_transpose through the C ABI on HOST buffers (pinned): upload of A, transpose, download of
    A' inside the timed region (csb200_transpose_host)."""
    import ctypes as C
    from csparse_cuda import _lib
    nnz = len(i)
    hp, hi, hx = pinned(torch, p), pinned(torch, i), pinned(torch, x)
    cp = torch.empty(m + 1, dtype=torch.int32).pin_memory()
    ci = torch.empty(max(nnz, 1), dtype=torch.int32).pin_memory()
    cx = torch.empty(max(nnz, 1), dtype=torch.float64).pin_memory()
    ptr = lambda t: C.c_void_p(t.data_ptr())
    call = lambda: _lib.check(_lib.lib().csb200_transpose_host(m, n, ptr(hp), ptr(hi), ptr(hx), ptr(cp), ptr(ci), ptr(cx)))
    call(); call()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    dt = (time.perf_counter() - t0) / steps
    from csparse_cuda import synth
    b = synth.transpose_bytes(m, n, nnz)
    return {"value": b / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 4 * (n + 1) + 12 * nnz,
            "d2h_bytes_per_step": 4 * (m + 1) + 12 * nnz, "ms_per_step": dt * 1e3, "steps": steps,
            "call": "csb200_transpose_host(m,n,Ap,Ai,Ax,Cp,Ci,Cx): pinned host buffers; uploads A, transposes, downloads A'"}


def e2e_multiply(torch, cc, m, n, p, i, x, nnzc, steps=3):
    """cs_multiply A*A through the C ABI on HOST buffers (pinned): upload of A, multiply, download of
    C inside the timed region (csb200_mat_upload + csb200_multiply + csb200_mat_download)."""
    import ctypes as C
    from csparse_cuda import _lib
    L = _lib.lib()
    nnz = len(i)
    hp, hi, hx = pinned(torch, p), pinned(torch, i), pinned(torch, x)
    cp = torch.empty(n + 1, dtype=torch.int32).pin_memory()
    ci = torch.empty(nnzc, dtype=torch.int32).pin_memory()
    cx = torch.empty(nnzc, dtype=torch.float64).pin_memory()
    ptr = lambda t: C.c_void_p(t.data_ptr())

    def call():
        hA, hC = C.c_void_p(), C.c_void_p()
        _lib.check(L.csb200_mat_upload(m, n, ptr(hp), ptr(hi), ptr(hx), 1, C.byref(hA)))
        _lib.check(L.csb200_multiply(hA, hA, C.byref(hC)))
        _lib.check(L.csb200_mat_download(hC, ptr(cp), ptr(ci), ptr(cx)))
        L.csb200_mat_free(hC)
        L.csb200_mat_free(hA)
    call(); call()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
    dt = (time.perf_counter() - t0) / steps
    assert int(cp[n]) == nnzc
    return {"value": nnzc / dt, "unit": "nnz(C)/s", "h2d_bytes_per_step": 4 * (n + 1) + 12 * nnz,
            "d2h_bytes_per_step": 4 * (n + 1) + 12 * nnzc, "ms_per_step": dt * 1e3, "steps": steps,
            "call": "csb200_mat_upload(A) + csb200_multiply(A, A) + csb200_mat_download(C): pinned host buffers, "
                    "validation of A, pattern classes / entry-major copy of A rebuilt every step (fresh handle)"}


def bench_gaxpy_lap2d(a, torch, dist, cc, synth, csd, world, rank, k, peak, peak_src, sampler):
    """Row-block sharded cs_gaxpy on the 5-point Laplacian.  N = 1: the k x k grid.
    N > 1 weak: a k x (k N) grid, one k x k slab per rank; strong: the k x k grid split by rows."""
    ky_total = k * world if a.scaling == "weak" else k
    n_global = k * ky_total
    bounds = csd.even_bounds(ky_total, world) * k          # split between grid lines
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    # the Laplacian is symmetric: rows r0..r1 of the CSR view == columns r0..r1 of the CSC
    m_, n_, p, i, x = synth.lap2d_cols(k, ky_total, r0, r1)
    nnz_local = len(i)
    nnz_global = 5 * n_global - 2 * k - 2 * ky_total
    alg_bytes_local = synth.gaxpy_bytes(r1 - r0, r1 - r0, nnz_local)
    alg_bytes_global = synth.gaxpy_bytes(n_global, n_global, nnz_global)

    x_own = torch.from_numpy(np.random.default_rng(rank).standard_normal(r1 - r0)).cuda()
    y_own = torch.from_numpy(np.random.default_rng(100 + rank).standard_normal(r1 - r0)).cuda()
    if world == 1:
        dA = cc.from_arrays(n_global, n_global, p, i, x)     # CSC; its CSR view is built once and cached
        dA.prepare_gaxpy()
        plan = dA.gaxpy_plan()
        xp, yp = x_own.data_ptr(), y_own.data_ptr()
        step = lambda: dA.gaxpy_dev(xp, yp)
        launches_per_step, exch = 1 if plan == "stream" else 2, 0
        mode = "single"
    else:
        blk = csd.RowBlock(r0, r1, p, i, x, int(i.min()), int(i.max()))
        sh = csd.ShardedGaxpy(blk, n_global, n_global, bounds, make_local=csd.cuda_make_local,
                              local_spmv=csd.cuda_local_spmv, device="cuda", fused=(a.halo == "fused"))
        plan = sh.handle.gaxpy_plan()
        xv = sh.own_view()                      # x lives inside the halo window: no per-step copy
        xv.copy_(x_own)
        x_own = xv
        step = lambda: sh.step(x_own, y_own)
        step()
        per = 1 if plan == "stream" else 2
        if sh.fused:
            launches_per_step = per if plan == "stream" else per + 2
            mode = "halo pulled over NVLink by the SpMV launch itself (peer-mapped windows, no collective)"
        else:
            launches_per_step = per * (1 + (sh.h_top is not None) + (sh.h_bot is not None)) if sh.split else per
            mode = sh.plan.mode + ("+overlap (NCCL batch_isend_irecv)" if sh.split else "")
        exch = sh.exchanged_bytes

    l0 = cc.launch_count()
    sampler.start()
    ms = device_timed(torch, dist, world, step, a.steps, a.warmup)
    launches = cc.launch_count() - l0 - launches_per_step * a.warmup
    value = alg_bytes_global * a.steps / (ms * 1e-3) / 1e9

    # kernel-only roofline of the dominant kernel: the local SpMV launch, timed alone
    if world == 1:
        kern = step
    else:
        xw = sh.x_window
        if sh.fused:     # one handle for the whole block: the same rows without the halo protocol
            kern = lambda: csd.cuda_local_spmv(sh.handle, xw, y_own)
        else:
            mid = y_own[sh.split[0]:sh.split[1]] if sh.split else y_own
            kern = lambda: csd.cuda_local_spmv(sh.handle, xw, mid)
            if sh.split:     # the interior block is the dominant launch
                nr = sh.split[1] - sh.split[0]
                alg_bytes_local = synth.gaxpy_bytes(nr, nr, sh.handle.nnz)
    kms = device_timed(torch, dist, 1, kern, a.steps, 2) / a.steps
    clocks = sampler.stop()
    achieved = alg_bytes_local / (kms * 1e-3) / 1e9
    # DRAM bytes per launch of the dominant kernel on this exact config, from the committed ncu
    # --set full capture of the current binary (profiles/traffic.json names the capture it came from)
    traffic = measured_traffic(f"k_spmv_tma lap2d {k}^2") if (world == 1 and plan == "stream") else None

    # e2e: the reference-facing call on HOST buffers (pinned), copies inside the timed region.
    # "e2e": csb200_gaxpy(handle, x, y) -- the step's inputs (x and the y it accumulates into)
    # travel H2D and the result y travels D2H every step; the matrix is the handle's resident
    # state, uploaded once like the cs object the reference keeps between calls.
    # "e2e_cold": csb200_gaxpy_host -- the one-shot form, matrix uploaded and CSR view rebuilt
    # every step as well.
    e2e, e2e_res = None, None
    if world == 1:
        hp, hi, hx = pinned(torch, p), pinned(torch, i), pinned(torch, x)
        hxv = pinned(torch, np.random.default_rng(0).standard_normal(n_global))
        hyv = pinned(torch, np.random.default_rng(1).standard_normal(n_global))
        cold = lambda: cc.gaxpy_host(n_global, n_global, hp.data_ptr(), hi.data_ptr(), hx.data_ptr(),
                                     hxv.data_ptr(), hyv.data_ptr())
        ksteps = max(3, min(a.steps, 10))
        cold(); cold()
        t0 = time.perf_counter()
        for _ in range(ksteps):
            cold()
        dt = (time.perf_counter() - t0) / ksteps
        h2d = hp.numel() * 4 + hi.numel() * 4 + hx.numel() * 8 + 8 * n_global * 2
        e2e_cold = {"value": alg_bytes_global / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 8 * n_global, "ms_per_step": dt * 1e3, "steps": ksteps,
               "call": "csb200_gaxpy_host(m,n,Ap,Ai,Ax,x,y): pinned host buffers; uploads the matrix, builds the CSR view, SpMV, downloads y"}
        import ctypes as C
        from csparse_cuda import _lib
        res = lambda: _lib.check(_lib.lib().csb200_gaxpy(dA._h, C.c_void_p(hxv.data_ptr()), C.c_void_p(hyv.data_ptr())))
        res(); res()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            res()
        dt = (time.perf_counter() - t0) / a.steps
        e2e = {"value": alg_bytes_global / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 16 * n_global,
               "d2h_bytes_per_step": 8 * n_global, "ms_per_step": dt * 1e3, "steps": a.steps,
               "call": "csb200_gaxpy(handle, x, y): pinned host x and y copied H2D and y copied back D2H every step; "
                       "the matrix handle (uploaded once, like the reference's cs object) stays in HBM"}
        e2e_res = e2e_cold
    else:
        # N > 1, host slices of x and y per rank (pinned, allocated after binding the process to the
        # CPUs next to its GPU).  Fused halo: csb200_gaxpy_halo -- the ends of x first, the neighbours'
        # lines pulled over NVLink, y in row chunks with duplex copies, as csb200_gaxpy does at N = 1.
        node = bind_numa(torch.cuda.current_device())
        hx_own, hy_own = x_own.cpu().pin_memory(), y_own.cpu().pin_memory()
        if sh.fused:
            # y accumulates in HBM between steps (as in a solver loop); the host gets the step's result.
            # All ranks of this box share two PCIe uplinks (~106 GB/s H2D in aggregate, measured at N = 8
            # with y travelling up as well: 20.2 ms per step), so the bytes per step are what counts.
            e2e_step = lambda: sh.step_host(hx_own, hy_own, y_own)
            call = ("per rank: csb200_gaxpy_halo(block, halo, x_slice, y_copy, d_y) on pinned host slices -- x H2D (ends first), "
                    "halo lines pulled from the neighbours' windows, y D2H in row chunks (duplex), y and the matrix block resident")
            h2d = 8 * (r1 - r0)
        else:
            def e2e_step():
                x_own.copy_(hx_own, non_blocking=True)
                sh.step(x_own, y_own)
                hy_own.copy_(y_own, non_blocking=True)
            call = "per rank: x slice H2D, halo exchange + local csb200_gaxpy_t_dev, y slice D2H (matrix block resident)"
            h2d = 8 * (r1 - r0)
        esteps = min(a.steps, 100)
        ems = device_timed(torch, dist, world, e2e_step, esteps, 2)
        e2e = {"value": alg_bytes_global * esteps / (ems * 1e-3) / 1e9, "unit": "GB/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8 * (r1 - r0), "ms_per_step": ems / esteps,
               "steps": esteps, "numa_node_of_rank0_buffers": node, "call": call}

    res = {
        "metric": "cs_gaxpy HBM GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"cs_gaxpy y+=A*x, 2-D 5-point Laplacian {k}x{ky_total} (n={n_global}, nnz={nnz_global}), "
                               f"{r1 - r0} rows/GPU, row-block sharded", "kernel": "k_spmv_tma" if plan == "stream" else f"k_spmv_{plan}",
                   "exchange": mode, "exchange_bytes_per_rank_step": exch,
                   "l2": "inputs (1.0 GB matrix per GPU) exceed the 126 MB L2; no flush needed",
                   "gflops": 2 * nnz_global * a.steps / (ms * 1e-3) / 1e9},
        "clocks": clocks, "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "kernel": "k_spmv_tma" if plan == "stream" else f"k_spmv_{plan}",
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_local, "kernel_ms": kms},
        "e2e": e2e,
    }
    if e2e_res:
        res["e2e_cold"] = e2e_res
    return res


def bench_gaxpy_single(a, torch, cc, synth, family, k, peak, peak_src, sampler):
    if family == "rmat":
        m, n, tp, ti, tx = synth.rmat_torch(k, 16)          # generated on the GPU (inputs only)
        nnz = int(ti.numel())
        dA = cc.from_device(m, n, tp.data_ptr(), ti.data_ptr(), tx.data_ptr())
        torch.cuda.synchronize()
        del tp, ti, tx
        torch.cuda.empty_cache()
    else:
        m, n, p, i, x = synth.lap2d(k)
        nnz = len(i)
        dA = cc.from_arrays(m, n, p, i, x)
    t0 = time.perf_counter()
    dA.prepare_gaxpy()
    cc.synchronize()
    t_csr = time.perf_counter() - t0
    plan = dA.gaxpy_plan()
    xv = torch.randn(n, dtype=torch.float64, device="cuda")
    yv = torch.randn(m, dtype=torch.float64, device="cuda")
    step = lambda: dA.gaxpy_dev(xv.data_ptr(), yv.data_ptr())
    l0 = cc.launch_count()
    sampler.start()
    ms = device_timed(torch, None, 1, step, a.steps, a.warmup)
    clocks = sampler.stop()
    per = 1 if plan == "stream" else 2
    launches = cc.launch_count() - l0 - per * a.warmup
    b = synth.gaxpy_bytes(m, n, nnz)
    value = b * a.steps / (ms * 1e-3) / 1e9
    return {"metric": "cs_gaxpy HBM GB/s", "value": value, "unit": "GB/s", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cs_gaxpy, R-MAT scale {k} ef 16 (n={n}, nnz={nnz})", "kernel": f"k_spmv_{plan}",
                       "csr_view_build_ms_cold": t_csr * 1e3,
                       "l2": "inputs exceed L2" if b > 2.5e8 else "inputs fit L2 (no flush): launch-bound parity config"},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": value, "peak": peak, "unit": "GB/s", "frac": value / peak,
                         "traffic": None, "kernel": f"k_spmv_{plan}", "peak_source": peak_src},
            "e2e": None}


def bench_gaxpy_rmat_dist(a, torch, dist, cc, synth, csd, world, rank, k, peak, peak_src, sampler):
    """cs_gaxpy on the R-MAT matrix over N GPUs (strong scaling): contiguous row blocks of the CSR
    view balanced by nonzeros, x owned in equal slices and all-gathered every step (NCCL), then the
    local merge-path SpMV.  y is row-owned: no reduction."""
    m, n, tp, ti, tx = synth.rmat_torch(k, 16)                 # same seed on every rank: the same matrix
    nnz = int(ti.numel())
    dA = cc.from_device(m, n, tp.data_ptr(), ti.data_ptr(), tx.data_ptr())
    torch.cuda.synchronize()
    del tp, ti, tx
    torch.cuda.empty_cache()
    dAT = cc.cs_transpose(dA, True)                            # CSC of A' == CSR view of A
    dA.free()
    p_ptr, _, _ = dAT.device_pointers()
    rowptr = csd._as_tensor(p_ptr, m + 1, torch.int32, "cuda").to(torch.int64)
    targets = torch.arange(world + 1, device="cuda", dtype=torch.int64) * nnz // world
    bounds = torch.searchsorted(rowptr, targets).clamp_(0, m)
    bounds[0], bounds[-1] = 0, m
    bounds = [int(v) for v in bounds.tolist()]
    r0, r1 = bounds[rank], bounds[rank + 1]
    blk = dAT.col_slice(r0, r1)                                # this rank's rows, all columns
    nnz_local = blk.nnz
    del rowptr
    assert n % world == 0
    xs = n // world
    x_full = torch.empty(n, dtype=torch.float64, device="cuda")
    x_own = torch.randn(xs, dtype=torch.float64, device="cuda")
    y_own = torch.randn(r1 - r0, dtype=torch.float64, device="cuda")
    # parity of the sharded step before anything is timed: the same rows from the whole matrix
    dist.all_gather_into_tensor(x_full, x_own)
    y_ref = torch.zeros(m, dtype=torch.float64, device="cuda")
    y_ref[r0:r1] = y_own
    dAT.gaxpy_t_dev(x_full.data_ptr(), y_ref.data_ptr())
    y_chk = y_own.clone()
    blk.gaxpy_t_dev(x_full.data_ptr(), y_chk.data_ptr())
    err = float(torch.linalg.norm(y_chk - y_ref[r0:r1]) / torch.linalg.norm(y_ref[r0:r1]))
    assert err <= 1e-12, f"sharded cs_gaxpy differs from the single-GPU result: {err:.3e}"
    del y_ref, y_chk
    dAT.free()

    def step():
        dist.all_gather_into_tensor(x_full, x_own)
        blk.gaxpy_t_dev(x_full.data_ptr(), y_own.data_ptr())
    step()
    l0 = cc.launch_count()
    sampler.start()
    ms = device_timed(torch, dist, world, step, a.steps, a.warmup)
    launches = (cc.launch_count() - l0) * a.steps // (a.steps + a.warmup)
    kern = lambda: blk.gaxpy_t_dev(x_full.data_ptr(), y_own.data_ptr())
    kms = device_timed(torch, dist, 1, kern, a.steps, 2) / a.steps
    clocks = sampler.stop()
    b_global = synth.gaxpy_bytes(m, n, nnz)
    b_local = synth.gaxpy_bytes(r1 - r0, n, nnz_local)
    value = b_global * a.steps / (ms * 1e-3) / 1e9
    hx, hy = x_own.cpu().pin_memory(), y_own.cpu().pin_memory()

    def e2e_step():
        x_own.copy_(hx, non_blocking=True)
        step()
        hy.copy_(y_own, non_blocking=True)
    ems = device_timed(torch, dist, world, e2e_step, min(a.steps, 200), 2)
    esteps = min(a.steps, 200)
    return {"metric": "cs_gaxpy HBM GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cs_gaxpy, R-MAT scale {k} ef 16 (n={n}, nnz={nnz}), row blocks balanced by nonzeros "
                                   f"({r1 - r0} rows / {nnz_local} nnz on rank 0), x all-gathered every step",
                       "kernel": "k_spmv_long + k_spmv_mid + k_spmv_short (rows binned by length) / k_spmv_tma, by row statistics of the block",
                       "exchange": "all-gather", "exchange_bytes_per_rank_step": 8 * (n - xs),
                       "l2": "inputs exceed L2"},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": b_local / (kms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": b_local / (kms * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "local SpMV of rank 0",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": b_local, "kernel_ms": kms},
            "e2e": {"value": b_global * esteps / (ems * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": 8 * xs,
                    "d2h_bytes_per_step": 8 * (r1 - r0), "ms_per_step": ems / esteps,
                    "call": "per rank: x slice H2D, all-gather of x, local csb200_gaxpy_t_dev, y slice D2H (row block resident)"}}


def bench_transpose(a, torch, cc, synth, k, peak, peak_src, sampler):
    m, n, p, i, x = synth.lap2d(k)
    nnz = len(i)
    dA = cc.from_arrays(m, n, p, i, x)
    hold = {}

    def step():
        hold["c"] = cc.cs_transpose(dA, True)
    l0 = cc.launch_count()
    sampler.start()
    ms = device_timed(torch, None, 1, step, a.steps, a.warmup)
    clocks = sampler.stop()
    total_l = cc.launch_count() - l0
    launches = total_l * a.steps // (a.steps + a.warmup)
    b = synth.transpose_bytes(m, n, nnz)
    value = b * a.steps / (ms * 1e-3) / 1e9
    return {"metric": "cs_transpose HBM GB/s", "value": value, "unit": "GB/s", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": a.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cs_transpose(values=True), lap2d {k}^2 (nnz={nnz})", "l2": "inputs exceed L2"},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": value, "peak": peak, "unit": "GB/s", "frac": value / peak,
                         "traffic": None, "kernel": "whole cs_transpose (hist+scan+scatter+fix)", "peak_source": peak_src},
            "e2e": e2e_transpose(torch, cc, m, n, p, i, x)}


def bench_multiply(a, torch, dist, cc, synth, csd, world, rank, k, peak, peak_src, sampler):
    """cs_multiply A*A on the 27-point stencil; N > 1: column blocks of B, A replicated, final gather."""
    m, n, p, i, x = synth.st27(k)
    nnz = len(i)
    dA = cc.from_arrays(m, n, p, i, x)
    bounds = csd.multiply_column_bounds(p, p, i, world)
    hold = {}
    # one result alive at a time: the previous product is released before the next is formed,
    # so the stream-ordered pool reuses its blocks instead of growing (a fresh GB costs ~100 ms)
    if world == 1:
        def step():
            hold.clear()
            hold["c"] = cc.cs_multiply(dA, dA)
    else:
        dBl = dA.col_slice(int(bounds[rank]), int(bounds[rank + 1]))     # this rank's columns of B, sliced once

        def step():
            hold.clear()
            hold["c"] = csd.sharded_multiply(dA, dA, bounds, rank, gather="root", device="cuda", dB_local=dBl)
    steps, warm = min(a.steps, 10), max(min(a.warmup, 3), 2)
    l0 = cc.launch_count()
    sampler.start()
    ms = device_timed(torch, dist, world, step, steps, warm)
    clocks = sampler.stop()
    ms_local = None
    if world > 1:
        # the same step with C left column-distributed (no final gather): the compute part alone
        def step_local():
            hold.clear()
            hold["c"] = csd.sharded_multiply(dA, dA, bounds, rank, gather=None, device="cuda", dB_local=dBl)
        ms_local = device_timed(torch, dist, world, step_local, steps, warm)
        step()                                     # leave a gathered product for the checks below
    launches = (cc.launch_count() - l0) * steps // (steps + warm)
    nnzc = (5 * k - 6) ** 3
    if world == 1:
        got = hold["c"].nnz
    elif rank == 0:
        got = int(hold["c"][1][1].numel())              # the gathered Ci on rank 0
    else:
        got = nnzc
    assert got == nnzc, (got, nnzc)
    value = nnzc * steps / (ms * 1e-3)
    b = synth.multiply_bytes(nnz, nnz, nnzc, n, n)
    gbs = b * steps / (ms * 1e-3) / 1e9
    return {"metric": "cs_multiply nnz(C)/s", "value": value, "unit": "nnz(C)/s", "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"cs_multiply A*A, 27-point stencil {k}^3 (n={n}, nnz={nnz}, nnz(C)={nnzc}), "
                                   f"column blocks of B, A replicated", "l2": "inputs + output exceed L2",
                       "gflops": 2 * cc.last_multiply_flops() * world * steps / (ms * 1e-3) / 1e9 if world == 1 else None,
                       "gather": "none (N = 1)" if world == 1 else "the whole product gathered to rank 0 (point-to-point pieces into their final offsets), inside the timed step",
                       "nnz(C)/s with C left column-distributed": None if ms_local is None else nnzc * steps / (ms_local * 1e-3),
                       "ms_per_step with C left column-distributed": None if ms_local is None else ms_local / steps},
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                         "traffic": None, "kernel": "whole cs_multiply (ub+bin+symbolic+scan+numeric)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": b},
            "e2e": e2e_multiply(torch, cc, m, n, p, i, x, nnzc) if world == 1 else None}


def extras(a, torch, cc, synth, peak):
    """Secondary device-resident numbers reported beside the headline at N = 1."""
    ex = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def timed(fn, warm, iters, batches=3):
        # ms per call: `batches` runs of `iters` back-to-back calls, one CUDA-event pair around each run
        # (host work of a call overlaps the previous call's kernels, as in a solver loop), median over
        # the runs.  These calls allocate their GB-sized results from the stream-ordered pool; a call
        # that makes the pool map fresh memory costs tens of ms once, which a single mean would report.
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        per_call = []
        for _ in range(batches):
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            per_call.append(e0.elapsed_time(e1) / iters)
        return float(np.median(per_call))

    try:
        m, n, p, i, x = synth.lap2d(4096)
        dA = cc.from_arrays(m, n, p, i, x)
        hold = {}
        def tr_paths(tag, dM, nb):
            # automatic choice (one-pass mirror path for these symmetric-pattern stencils) and the
            # general two-hop bucket sort on the same matrix
            for path in (None, "bucket"):
                cc.force_transpose_path(path)
                try:
                    # 12 warm-up calls: the bucket path's temporaries (two nnz-sized intermediates next to
                    # the result) take the stream-ordered pool ~15 calls to stop mapping fresh memory
                    # (tools/diag_pool.py: 7.1, 4.6, 2.0 ms per call over the first three runs of five,
                    # 1.27 ms from then on in every mode)
                    t = timed(lambda: hold.__setitem__("c", cc.cs_transpose(dM, True)), 12, 5)
                    took = cc.last_transpose_path()
                finally:
                    cc.force_transpose_path(None)
                ex[f"cs_transpose {tag}" + ("" if path is None else " (bucket path)")] = {
                    "ms": t, "GB/s": nb / t / 1e6, "frac_of_peak": nb / t / 1e6 / peak, "path": took}
                hold.clear()

        tr_paths("lap2d 4096^2", dA, synth.transpose_bytes(m, n, len(i)))
        hold.clear(); dA.free()
        ex["cs_transpose lap2d 4096^2 e2e (host buffers through the C ABI)"] = e2e_transpose(torch, cc, m, n, p, i, x)
        m, n, p, i, x = synth.st27(128)
        dA = cc.from_arrays(m, n, p, i, x)
        def mul():
            hold.clear()
            hold["c"] = cc.cs_multiply(dA, dA)
        ms = timed(mul, 2, 5)
        nnzc = hold["c"].nnz
        b = synth.multiply_bytes(len(i), len(i), nnzc, n, n)
        ex["cs_multiply st27 128^3 A*A"] = {"ms": ms, "nnz(C)/s": nnzc / ms * 1e3, "GB/s": b / ms / 1e6,
                                            "frac_of_peak": b / ms / 1e6 / peak, "nnzC": nnzc,
                                            "GFLOP/s": 2 * cc.last_multiply_flops() / ms / 1e6}
        hold.clear()
        # the same product through the general kernels (what a matrix without translation-invariant columns gets)
        cc.force_multiply_path("no_templates")
        try:
            ms = timed(mul, 2, 3)
        finally:
            cc.force_multiply_path(None)
        ex["cs_multiply st27 128^3 A*A (general kernels, pattern classes switched off)"] = {
            "ms": ms, "nnz(C)/s": nnzc / ms * 1e3, "GB/s": b / ms / 1e6, "frac_of_peak": b / ms / 1e6 / peak}
        hold.clear()
        tr_paths("st27 128^3", dA, synth.transpose_bytes(m, n, len(i)))
        hold.clear(); dA.free()
        ex["cs_multiply st27 128^3 A*A e2e (host buffers through the C ABI)"] = e2e_multiply(torch, cc, m, n, p, i, x, nnzc)
        # cs_cumsum (csparse.py:767-784) on 2^24 counts: one API call = memset + launch + the synchronize that
        # returns the total (written by the kernel into pinned host memory)
        import ctypes as C
        from csparse_cuda import _lib
        nc = 1 << 24
        # c <- p[0..n-1] is part of the contract, so repeated calls on one array square the counts: zeros are
        # the only input that stays put (the kernel's work does not depend on the values)
        cvec = torch.zeros(nc, dtype=torch.int32, device="cuda")
        pvec = torch.empty(nc + 1, dtype=torch.int32, device="cuda")
        tot = C.c_int64()
        ms = timed(lambda: _lib.check(_lib.lib().csb200_cumsum_dev(C.c_void_p(pvec.data_ptr()), C.c_void_p(cvec.data_ptr()),
                                                                   nc, C.byref(tot))), 3, 10)
        b = synth.cumsum_bytes(nc)
        ex["cs_cumsum n = 2^24 (device arrays, total returned to the host)"] = {"ms": ms, "GB/s": b / ms / 1e6,
                                                                              "frac_of_peak": b / ms / 1e6 / peak}
        del cvec, pvec
        m, n, tp, ti, tx = synth.rmat_torch(24, 16)
        nnz = int(ti.numel())
        dA = cc.from_device(m, n, tp.data_ptr(), ti.data_ptr(), tx.data_ptr())
        torch.cuda.synchronize()
        del tp, ti, tx
        torch.cuda.empty_cache()
        def tr():
            hold.clear()                     # one result alive at a time (3.2 GB each)
            hold["c"] = cc.cs_transpose(dA, True)
        ms = timed(tr, 2, 3)
        b = synth.transpose_bytes(m, n, nnz)
        ex["cs_transpose rmat 2^24 (radix path)"] = {"ms": ms, "GB/s": b / ms / 1e6, "frac_of_peak": b / ms / 1e6 / peak}
        hold.clear()
        dA.prepare_gaxpy()
        xv = torch.randn(n, dtype=torch.float64, device="cuda")
        yv = torch.randn(m, dtype=torch.float64, device="cuda")
        ms = timed(lambda: dA.gaxpy_dev(xv.data_ptr(), yv.data_ptr()), 3, 20)
        b = synth.gaxpy_bytes(m, n, nnz)
        ex[f"cs_gaxpy rmat 2^24 ({dA.gaxpy_plan()} plan)"] = {"ms": ms, "GB/s": b / ms / 1e6, "frac_of_peak": b / ms / 1e6 / peak,
                                                 "nnz": nnz, "GFLOP/s": 2 * nnz / ms / 1e6}
        dA.free()
        # the rows either side of the hot path, lap2d 4096^2, device-resident
        m, n, p, i, x = synth.lap2d(4096)
        nnz = len(i)
        dA = cc.from_arrays(m, n, p, i, x)
        dAT = cc.cs_transpose(dA, True)

        def add():
            hold.clear()
            hold["c"] = cc.cs_add(dA, dAT, 1.0, 1.0)
        ms = timed(add, 2, 5)
        b = 12 * (2 * nnz + hold["c"].nnz) + 12 * (n + 1)
        ex["cs_add A+A' lap2d 4096^2"] = {"ms": ms, "GB/s": b / ms / 1e6, "frac_of_peak": b / ms / 1e6 / peak}
        hold.clear(); dAT.free()
        cols = torch.repeat_interleave(torch.arange(n, dtype=torch.int32, device="cuda"),
                                       torch.from_numpy(np.diff(p).astype(np.int64)).cuda())
        perm = torch.randperm(nnz, device="cuda")
        tj, ti, tx = cols[perm].contiguous(), torch.from_numpy(i).cuda()[perm].contiguous(), \
            torch.from_numpy(x).cuda()[perm].contiguous()
        del cols, perm
        import ctypes as C
        from csparse_cuda import _lib

        def compress():
            hold.clear()
            out = C.c_void_p()
            _lib.check(_lib.lib().csb200_compress_dev(m, n, nnz, C.c_void_p(ti.data_ptr()), C.c_void_p(tj.data_ptr()),
                                                      C.c_void_p(tx.data_ptr()), C.byref(out)))
            hold["c"] = cc.DeviceMatrix(out.value)
        ms = timed(compress, 2, 5)
        b = 32 * nnz + 4 * (n + 1)
        ex["cs_compress lap2d 4096^2 (shuffled triplets, radix path)"] = {"ms": ms, "GB/s": b / ms / 1e6,
                                                                         "frac_of_peak": b / ms / 1e6 / peak}
        hold.clear(); dA.free()
        del ti, tj, tx
        # BASELINE.json configs 1 and 2: the reference's own test matrices (tests/golden holds them), the
        # CSparseTest1 operations cs_transpose and cs_multiply A*A' -- device handles, host `cs` objects
        # backed by numpy arrays (upload, kernels, download: what a caller of the module pays), and the
        # C port of the reference on the same inputs (the Python reference is ~100x slower than the port)
        from oracle import oracle as orc
        orc.build()
        import json as _json

        def wall(fn, warm, iters):
            for _ in range(warm):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(iters):
                t0 = time.perf_counter()
                fn()
                torch.cuda.synchronize()
                ts.append((time.perf_counter() - t0) * 1e3)
            return float(np.median(ts))

        for name in ("bcsstk01", "bcsstk16"):
            z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tests", "golden", name + ".npz"))
            meta = _json.loads(str(z["meta"]))
            mm, nn = meta["A"]["m"], meta["A"]["n"]
            zp, zi, zx = z["A_p"], z["A_i"], z["A_x"]
            dS = cc.from_arrays(mm, nn, zp, zi, zx)
            dST = cc.cs_transpose(dS, True)
            S = cc.cs()
            S.m, S.n, S.nzmax, S.nz, S.p, S.i, S.x = mm, nn, len(zi), -1, zp.copy(), zi.copy(), zx.copy()
            ST = cc.cs_transpose(S, True)
            oS = orc.csc(mm, nn, zp, zi, zx)
            t0 = time.perf_counter(); oT = orc.cs_transpose(oS, True); t_ot = (time.perf_counter() - t0) * 1e3
            t0 = time.perf_counter(); oC = orc.cs_multiply(oS, oT); t_oc = (time.perf_counter() - t0) * 1e3
            ex[f"{name} (BASELINE config {1 if name == 'bcsstk01' else 2}): cs_transpose, cs_multiply A*A'"] = {
                "nnz": int(len(zi)), "nnzC": int(oC.nnz),
                "transpose ms: device handle": wall(lambda: hold.__setitem__("t", cc.cs_transpose(dS, True)), 3, 20),
                "transpose ms: host cs (numpy-backed, up + kernels + down)": wall(lambda: cc.cs_transpose(S, True), 2, 10),
                "transpose ms: C port of the reference, 1 core": t_ot,
                "multiply ms: device handles": wall(lambda: hold.__setitem__("c", cc.cs_multiply(dS, dST)), 3, 20),
                "multiply ms: host cs (numpy-backed, up + kernels + down)": wall(lambda: cc.cs_multiply(S, ST), 2, 10),
                "multiply ms: C port of the reference, 1 core": t_oc}
            hold.clear(); dS.free(); dST.free()
        # CPU baseline of the second headline metric: oracle cs_multiply on a bounded sample
        ms_, ns_, ps_, is_, xs_ = synth.st27(64)
        Ao = orc.csc(ms_, ns_, ps_, is_, xs_)
        t0 = time.perf_counter()
        Co = orc.cs_multiply(Ao, Ao)
        dt = time.perf_counter() - t0
        ex["cpu_baseline cs_multiply (oracle port, 1 core, st27 64^3 sample: same per-column work as 128^3, rate is flat in n)"] = {
            "s": dt, "nnz(C)/s": Co.nnz / dt, "nnzC": Co.nnz}
    except Exception as e:  # secondary numbers never sink the headline
        ex["error"] = repr(e)
    return ex


if __name__ == "__main__":
    main()
