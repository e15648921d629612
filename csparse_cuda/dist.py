"""Multi-GPU sharding of the two paths that shard, one process per GPU,
``torch.distributed`` for the plumbing (NCCL over NVLink/NVSwitch on the B200 box,
gloo in the CPU tests).

cs_gaxpy     rows are independent: each rank owns a contiguous block of rows of the
             CSR view (balanced by nnz) and the matching slices of x and y.  Per
             step the rank fetches the x entries its block references --
             * halo mode  (banded matrices): only from the two neighbouring ranks,
               a few KB each way, as batched P2P isend/irecv;
             * gather mode (general matrices): an all-gather of x --
             then runs the local row-block SpMV.  y needs no reduction.
cs_multiply  columns of C are independent: each rank owns a block of columns of B
             (balanced by multiply-adds), A is replicated, every rank runs
             symbolic + scan + numeric on its block; the only communication is the
             final gather of the per-block results.
cs_transpose / cs_cumsum do not shard ("replicas only").

The partitioning and exchange logic is device-agnostic; the local kernels are
injected (``local_spmv``) so that the world_size-2 gloo tests exercise exactly the
code the NCCL path runs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np


# ---- partitioning (pure numpy) --------------------------------------------------

def balanced_bounds(weights_prefix: np.ndarray, parts: int) -> np.ndarray:
    """Split items 0..n-1 into `parts` contiguous blocks of nearly equal weight.

    ``weights_prefix`` is the inclusive-exclusive prefix array of length n+1
    (e.g. a row-pointer array).  Returns ``bounds`` of length parts+1 with
    bounds[0] = 0, bounds[parts] = n, non-decreasing.
    """
    n = len(weights_prefix) - 1
    total = int(weights_prefix[n])
    targets = (np.arange(1, parts, dtype=np.float64) * total / parts)
    inner = np.searchsorted(weights_prefix, targets, side="left").astype(np.int64)
    inner = np.clip(inner, 0, n)
    lower = np.clip(inner - 1, 0, n)                      # the cut just before may be closer
    closer = np.abs(weights_prefix[lower] - targets) < np.abs(weights_prefix[inner] - targets)
    inner = np.where(closer, lower, inner)
    b = np.concatenate(([0], inner, [n])).astype(np.int64)
    return np.maximum.accumulate(b)


def even_bounds(n: int, parts: int) -> np.ndarray:
    return (np.arange(parts + 1, dtype=np.int64) * n) // parts


@dataclass
class RowBlock:
    """Rows [r0, r1) of a CSR view, with the column window it references."""
    r0: int
    r1: int
    rowptr: np.ndarray      # int32, r1-r0+1, rebased to 0
    col: np.ndarray         # int32 GLOBAL column ids
    val: np.ndarray         # float64
    cmin: int
    cmax: int               # inclusive; cmin > cmax when the block is empty


def csr_row_block(rowptr: np.ndarray, col: np.ndarray, val: np.ndarray, r0: int, r1: int) -> RowBlock:
    b, e = int(rowptr[r0]), int(rowptr[r1])
    c = np.ascontiguousarray(col[b:e], dtype=np.int32)
    rp = (rowptr[r0:r1 + 1].astype(np.int64) - b).astype(np.int32)
    cmin, cmax = (int(c.min()), int(c.max())) if e > b else (0, -1)
    return RowBlock(r0, r1, rp, c, np.ascontiguousarray(val[b:e], dtype=np.float64), cmin, cmax)


@dataclass
class ExchangePlan:
    mode: str               # "halo" | "gather"
    x_bounds: np.ndarray    # ownership of x: rank g owns [x_bounds[g], x_bounds[g+1])
    win_lo: int             # local x window = global columns [win_lo, win_hi)
    win_hi: int
    lo_need: List[int]      # per rank: entries needed from the rank below
    hi_need: List[int]      # per rank: entries needed from the rank above


def plan_exchange(rank: int, world: int, x_bounds: np.ndarray, windows: Sequence[Sequence[int]],
                  n_global: int, force_gather: bool = False) -> ExchangePlan:
    """``windows[g] = (cmin, cmax)`` of every rank (all-gathered).  Halo mode needs every
    rank's window to stay inside its own slice of x plus its two neighbours' slices."""
    lo_need, hi_need, halo_ok = [], [], True
    for g in range(world):
        c0, c1 = int(x_bounds[g]), int(x_bounds[g + 1])
        cmin, cmax = windows[g]
        lo = max(0, c0 - cmin) if cmax >= cmin else 0
        hi = max(0, cmax - (c1 - 1)) if cmax >= cmin else 0
        lo_room = c0 - int(x_bounds[g - 1]) if g > 0 else 0
        hi_room = int(x_bounds[g + 2]) - c1 if g + 1 < world else 0
        if lo > lo_room or hi > hi_room:
            halo_ok = False
        lo_need.append(lo)
        hi_need.append(hi)
    if world == 1:
        return ExchangePlan("halo", x_bounds, 0, n_global, [0], [0])
    if not halo_ok or force_gather:
        return ExchangePlan("gather", x_bounds, 0, n_global, lo_need, hi_need)
    c0, c1 = int(x_bounds[rank]), int(x_bounds[rank + 1])
    return ExchangePlan("halo", x_bounds, c0 - lo_need[rank], c1 + hi_need[rank], lo_need, hi_need)


class ShardedGaxpy:
    """y_local += A[r0:r1, :] * x for this rank's row block; x and y are distributed.

    Parameters
    ----------
    block : RowBlock            this rank's rows of the CSR view (global column ids)
    m_global, n_global : int    shape of A
    row_bounds : array(G+1)     row ownership (y follows it)
    x_bounds : array(G+1)       ownership of x (defaults to row_bounds when A is square)
    make_local : callable(rowptr, col_local, val, ncols_local) -> handle
    local_spmv : callable(handle, x_window_tensor, y_local_tensor) -> None   (y += ...)
    device : torch device of the vectors
    """

    def __init__(self, block: RowBlock, m_global: int, n_global: int, row_bounds, x_bounds=None,
                 make_local: Callable = None, local_spmv: Callable = None, device="cpu", group=None,
                 force_gather: bool = False, fused: bool = False):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.block, self.device = block, device
        self.m_global, self.n_global = m_global, n_global
        self.row_bounds = np.asarray(row_bounds, dtype=np.int64)
        if x_bounds is None:
            x_bounds = self.row_bounds if m_global == n_global else even_bounds(n_global, self.world)
        self.x_bounds = np.asarray(x_bounds, dtype=np.int64)
        # every rank learns every rank's column window
        mine = torch.tensor([block.cmin, block.cmax], dtype=torch.int64)
        if self.world > 1:
            allw = [torch.zeros(2, dtype=torch.int64) for _ in range(self.world)]
            if dist.get_backend(group) == "nccl":
                allw_d = [w.to(device) for w in allw]
                dist.all_gather(allw_d, mine.to(device), group=group)
                allw = [w.cpu() for w in allw_d]
            else:
                dist.all_gather(allw, mine, group=group)
            windows = [(int(w[0]), int(w[1])) for w in allw]
        else:
            windows = [(block.cmin, block.cmax)]
        self.plan = plan_exchange(self.rank, self.world, self.x_bounds, windows, n_global, force_gather)
        pl = self.plan
        self.c0, self.c1 = int(self.x_bounds[self.rank]), int(self.x_bounds[self.rank + 1])
        # fused: the halos are pulled over NVLink inside the one SpMV launch (csb200_gaxpy_halo_dev)
        # from peer-mapped windows; otherwise batched NCCL / gloo send-recv fills a torch tensor
        self.fused = bool(fused) and pl.mode == "halo" and self.world > 1
        self.halo = None
        if self.fused:
            self.halo = FusedHalo(pl.win_hi - pl.win_lo, device)
            self.x_window = self.halo.window
        else:
            self.x_window = torch.zeros(pl.win_hi - pl.win_lo, dtype=torch.float64, device=device)
        self.own = slice(self.c0 - pl.win_lo, self.c1 - pl.win_lo)       # my slice inside the window
        col_local = (block.col.astype(np.int64) - pl.win_lo).astype(np.int32)
        self.local_spmv = local_spmv
        self.exchanged_bytes = 0
        # Rows that read halo entries sit (for banded matrices) at the two ends of the block:
        # split it into top / interior / bottom so the interior SpMV overlaps the exchange.
        nrows = block.r1 - block.r0
        self.split = None
        if pl.mode == "halo" and self.world > 1 and nrows > 0 and len(col_local):
            rp = block.rowptr.astype(np.int64)
            nonempty = rp[1:] > rp[:-1]
            starts = rp[:-1][nonempty]
            rmin = np.full(nrows, np.iinfo(np.int64).max)
            rmax = np.full(nrows, -1, dtype=np.int64)
            rmin[nonempty] = np.minimum.reduceat(block.col.astype(np.int64), starts)
            rmax[nonempty] = np.maximum.reduceat(block.col.astype(np.int64), starts)
            lo_rows = np.flatnonzero(rmin < self.c0)
            hi_rows = np.flatnonzero(rmax >= self.c1)
            top = int(lo_rows.max()) + 1 if len(lo_rows) else 0
            bot = int(hi_rows.min()) if len(hi_rows) else nrows
            if top < bot and (len(lo_rows) == 0 or lo_rows.max() < bot) and (len(hi_rows) == 0 or hi_rows.min() >= top):
                self.split = (top, bot)
        ncl = pl.win_hi - pl.win_lo
        if self.fused:
            # one handle for the whole block; the kernel runs the rows [top, bot) first and the rows
            # that read halo entries last (all of them when the block does not split)
            top, bot = self.split if self.split else (nrows, nrows)
            self.top_rows, self.bot_rows = top, nrows - bot
            self.handle = make_local(block.rowptr, col_local, block.val, ncl)
            self.h_top = self.h_bot = None
            self.halo.connect(self, group)
            self.exchanged_bytes = 8 * (pl.lo_need[self.rank] + pl.hi_need[self.rank])
        elif self.split:
            top, bot = self.split
            rp = block.rowptr

            def sub(a, b):
                if b <= a:
                    return None
                return make_local((rp[a:b + 1] - rp[a]).astype(np.int32), col_local[rp[a]:rp[b]],
                                  block.val[rp[a]:rp[b]], ncl)
            self.h_top, self.handle, self.h_bot = sub(0, top), sub(top, bot), sub(bot, nrows)
        else:
            self.handle = make_local(block.rowptr, col_local, block.val, ncl)

    def own_view(self):
        """This rank's slice of x inside the local window.  Writing x here (instead of passing a
        separate tensor to step) saves a device copy per step."""
        return self.x_window[self.own]

    def _place(self, x_own):
        view = self.x_window[self.own]
        if x_own.data_ptr() != view.data_ptr():
            view.copy_(x_own)

    # -- the per-step exchange of x --------------------------------------------------
    def exchange(self, x_own):
        """Fill the local x window from the distributed x (each rank passes its own slice)."""
        torch, dist, pl = self.torch, self.dist, self.plan
        xw = self.x_window
        if self.world == 1:
            self._place(x_own)
            return xw
        if pl.mode == "gather":
            sizes = [int(pl.x_bounds[g + 1] - pl.x_bounds[g]) for g in range(self.world)]
            if len(set(sizes)) == 1:
                dist.all_gather_into_tensor(xw, x_own.contiguous(), group=self.group)
            else:
                # uneven ownership: gather equal-sized padded slices, then place them
                mx = max(sizes)
                if getattr(self, "_pad", None) is None:
                    self._pad = torch.zeros(self.world * mx, dtype=torch.float64, device=self.device)
                    self._mine = torch.zeros(mx, dtype=torch.float64, device=self.device)
                self._mine[: sizes[self.rank]] = x_own
                dist.all_gather_into_tensor(self._pad, self._mine, group=self.group)
                for g in range(self.world):
                    xw[int(pl.x_bounds[g]):int(pl.x_bounds[g + 1])] = self._pad[g * mx: g * mx + sizes[g]]
            self.exchanged_bytes = 8 * (self.n_global - sizes[self.rank])
            return xw
        for w in self._exchange_start(x_own):
            w.wait()
        return xw

    def _exchange_start(self, x_own):
        """Halo mode: place the own slice, post the neighbour sends / receives, return the works."""
        dist, pl, xw = self.dist, self.plan, self.x_window
        self._place(x_own)
        r, ops, nbytes = self.rank, [], 0
        own = xw[self.own]
        if r > 0:
            give = pl.hi_need[r - 1]           # the rank below needs the head of my slice
            if give:
                ops.append(dist.P2POp(dist.isend, own[:give].contiguous(), r - 1, group=self.group))
            if pl.lo_need[r]:
                ops.append(dist.P2POp(dist.irecv, xw[: pl.lo_need[r]], r - 1, group=self.group))
                nbytes += 8 * pl.lo_need[r]
        if r + 1 < self.world:
            give = pl.lo_need[r + 1]           # the rank above needs the tail of my slice
            if give:
                ops.append(dist.P2POp(dist.isend, own[own.numel() - give:].contiguous(), r + 1, group=self.group))
            if pl.hi_need[r]:
                ops.append(dist.P2POp(dist.irecv, xw[xw.numel() - pl.hi_need[r]:], r + 1, group=self.group))
                nbytes += 8 * pl.hi_need[r]
        self.exchanged_bytes = nbytes
        return dist.batch_isend_irecv(ops) if ops else []

    def step_host(self, x_own_host, y_own_host, y_own_dev=None):
        """One distributed cs_gaxpy on pinned HOST tensors (fused mode): x slice up, halo pull, row
        chunks of y up / SpMV / down with duplex copies (csb200_gaxpy_halo); y_own_host += A_block * x.
        With ``y_own_dev`` y accumulates in that device tensor and y_own_host receives a copy."""
        if not self.fused:
            raise RuntimeError("step_host needs the fused halo exchange")
        pl, r = self.plan, self.rank
        edge_lo = pl.hi_need[r - 1] if r > 0 else 0              # the rank below reads the head of my slice
        edge_hi = pl.lo_need[r + 1] if r + 1 < self.world else 0
        self.halo.step_host(self.handle, x_own_host.data_ptr(), self.own.start, self.own.stop - self.own.start,
                            edge_lo, edge_hi, y_own_host.data_ptr() if y_own_host is not None else 0,
                            y_own_dev.data_ptr() if y_own_dev is not None else 0)
        return y_own_host

    def step(self, x_own, y_own):
        """One distributed cs_gaxpy: exchange x, then y_own += A_block * x_window.  In halo
        mode the interior rows run while the halos are in flight."""
        if self.fused:
            self._place(x_own)
            self.halo.step(self.handle, y_own, self.top_rows, self.bot_rows)
            return y_own
        if self.split:
            top, bot = self.split
            works = self._exchange_start(x_own)
            self.local_spmv(self.handle, self.x_window, y_own[top:bot])
            for w in works:
                w.wait()
            if self.h_top is not None:
                self.local_spmv(self.h_top, self.x_window, y_own[:top])
            if self.h_bot is not None:
                self.local_spmv(self.h_bot, self.x_window, y_own[bot:])
            return y_own
        xw = self.exchange(x_own)
        self.local_spmv(self.handle, xw, y_own)
        return y_own


# ---- fused halo exchange: peer-mapped x windows (libcsparse_b200.so, csb200_halo_*) -------------

class FusedHalo:
    """The rank's x window in IPC-exportable device memory plus the flag block the persistent SpMV
    kernel uses to pull the neighbours' halo lines over NVLink (csparse_cuda/csrc/spmv.cu).  The
    128-byte IPC handles are exchanged once with torch.distributed; per step there is no
    collective and no extra launch."""

    def __init__(self, count: int, device):
        import ctypes as C
        import torch
        from . import _lib
        self._lib, self._C = _lib, C
        h = C.c_void_p()
        _lib.check(_lib.lib().csb200_halo_create(int(count), C.byref(h)), "halo_create")
        self._h = h
        w = C.c_void_p()
        _lib.check(_lib.lib().csb200_halo_window(self._h, C.byref(w)), "halo_window")
        self.window = _as_tensor(w.value, max(int(count), 1), torch.float64, device)[: int(count)]
        self.count = int(count)

    def connect(self, sh: "ShardedGaxpy", group=None):
        """All ranks exchange (IPC handles, offset and length of the own slice inside the window);
        each then maps the two neighbours' windows."""
        import torch
        C, _lib, dist = self._C, self._lib, sh.dist
        buf = (C.c_ubyte * 128)()
        _lib.check(_lib.lib().csb200_halo_export(self._h, buf), "halo_export")
        mine = torch.zeros(128 + 16, dtype=torch.uint8)
        mine[:128] = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
        own_off, own_len = sh.own.start, sh.own.stop - sh.own.start
        mine[128:] = torch.frombuffer(bytearray(np.array([own_off, own_len], dtype=np.int64).tobytes()), dtype=torch.uint8)
        dev = sh.device
        allb = [torch.zeros_like(mine).to(dev) for _ in range(sh.world)]
        dist.all_gather(allb, mine.to(dev), group=group)
        allb = [b.cpu().numpy() for b in allb]
        pl, r = sh.plan, sh.rank
        for side, nb in ((0, r - 1), (1, r + 1)):
            if nb < 0 or nb >= sh.world:
                continue
            need = pl.lo_need[r] if side == 0 else pl.hi_need[r]
            if need == 0:
                continue
            off, length = (int(v) for v in np.frombuffer(allb[nb][128:].tobytes(), dtype=np.int64))
            peer_first = off + length - need if side == 0 else off          # its tail / its head
            local_first = 0 if side == 0 else self.count - need
            handles = (C.c_ubyte * 128).from_buffer_copy(allb[nb][:128].tobytes())
            _lib.check(_lib.lib().csb200_halo_connect(self._h, side, handles, peer_first, need, local_first),
                       "halo_connect")
        dist.barrier(group=group)

    def step(self, handle, y_own, top_rows: int, bot_rows: int):
        C, _lib = self._C, self._lib
        _lib.check(_lib.lib().csb200_gaxpy_halo_dev(handle._h, self._h, C.c_void_p(y_own.data_ptr()),
                                                    int(top_rows), int(bot_rows)), "gaxpy_halo_dev")

    def step_host(self, handle, x_own_ptr: int, own_off: int, own_len: int, edge_lo: int, edge_hi: int, y_ptr: int,
                  d_y_ptr: int = 0):
        """The step on HOST slices (raw addresses of pinned buffers): csb200_gaxpy_halo.  d_y_ptr: y
        accumulates in that device vector; the host slice at y_ptr (0 = none) gets a copy."""
        C, _lib = self._C, self._lib
        _lib.check(_lib.lib().csb200_gaxpy_halo(handle._h, self._h, C.c_void_p(x_own_ptr), int(own_off), int(own_len),
                                                int(edge_lo), int(edge_hi), C.c_void_p(y_ptr or None),
                                                C.c_void_p(d_y_ptr or None)), "gaxpy_halo")

    def timed_out(self) -> bool:
        v = self._C.c_int()
        self._lib.check(self._lib.lib().csb200_halo_status(self._h, self._C.byref(v)), "halo_status")
        return bool(v.value)

    def free(self):
        if self._h:
            self._lib.lib().csb200_halo_free(self._h)
            self._h = None


# ---- CUDA bindings for the local kernels -----------------------------------------------

def cuda_make_local(rowptr, col_local, val, ncols_local):
    """Upload a row block as the CSC of its transpose (= CSR view), ready for gaxpy_t_dev."""
    import csparse_cuda as cc
    h = cc.from_arrays(ncols_local, len(rowptr) - 1, rowptr, col_local, val, validate=True)
    return h


def cuda_local_spmv(handle, x_window, y_own):
    handle.gaxpy_t_dev(x_window.data_ptr(), y_own.data_ptr())


# ---- column-sharded cs_multiply ---------------------------------------------------------

def multiply_column_bounds(Ap: np.ndarray, Bp: np.ndarray, Bi: np.ndarray, parts: int) -> np.ndarray:
    """Column blocks of B with nearly equal multiply-add counts
    (flops[j] = sum over k in B(:,j) of nnz(A(:,k)))."""
    lens = np.diff(Ap).astype(np.int64)
    per_entry = lens[Bi[: int(Bp[-1])]]
    pref = np.zeros(len(per_entry) + 1, dtype=np.int64)
    np.cumsum(per_entry, out=pref[1:])
    col_prefix = pref[Bp.astype(np.int64)]          # flops before column j
    return balanced_bounds(col_prefix, parts)


def gather_columns(cp, ci, cx, bounds, rank: int, world: int, mode: str = "root", group=None, device="cpu"):
    """The final gather of a column-sharded product ("no collective beyond the final
    gather").  Every rank holds its block C(:, J_rank) as (cp rebased to 0, ci, cx or None);
    returns (Cp, Ci, Cx) of the whole matrix on rank 0 (mode "root"; None elsewhere) or on every
    rank (mode "all").  The only small collective is the all-gather of the world nnz counts; the
    pieces then travel ONCE, point to point, straight into their final offsets of the destination
    arrays -- no padding, no staging copy (NCCL send/recv over NVLink on the GPU box, gloo here).
    """
    import torch
    import torch.distributed as dist
    if mode not in ("root", "all"):
        raise ValueError("gather mode must be 'root' or 'all'")
    nnz_local = int(ci.numel())
    n_total = int(bounds[-1])
    has_x = cx is not None
    if world == 1:
        return cp.clone(), ci.clone(), (cx.clone() if has_x else None)
    cnt = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(cnt, torch.tensor([nnz_local], dtype=torch.int64, device=device), group=group)
    nnz_all = [int(v) for v in cnt.tolist()]
    offs = np.concatenate(([0], np.cumsum(nnz_all))).astype(np.int64)
    if int(offs[-1]) > 0x7FFFFFFF:
        raise OverflowError("gathered nnz(C) does not fit int32")
    cp_glob = (cp[:-1] + int(offs[rank])).contiguous()          # my column pointers in global numbering
    receivers = [0] if mode == "root" else list(range(world))
    Cp = Ci = Cx = None
    if rank in receivers:
        Cp = torch.empty(n_total + 1, dtype=torch.int32, device=device)
        Ci = torch.empty(int(offs[-1]), dtype=torch.int32, device=device)
        Cx = torch.empty(int(offs[-1]), dtype=torch.float64, device=device) if has_x else None
        Cp[n_total] = int(offs[-1])
    ops = []
    for dst in receivers:
        if dst == rank:
            for g in range(world):
                j0, j1, o0, o1 = int(bounds[g]), int(bounds[g + 1]), int(offs[g]), int(offs[g + 1])
                if g == rank:
                    Cp[j0:j1] = cp_glob
                    Ci[o0:o1] = ci
                    if has_x:
                        Cx[o0:o1] = cx
                    continue
                if j1 > j0:
                    ops.append(dist.P2POp(dist.irecv, Cp[j0:j1], g, group=group))
                if o1 > o0:
                    ops.append(dist.P2POp(dist.irecv, Ci[o0:o1], g, group=group))
                    if has_x:
                        ops.append(dist.P2POp(dist.irecv, Cx[o0:o1], g, group=group))
        else:
            if cp_glob.numel():
                ops.append(dist.P2POp(dist.isend, cp_glob, dst, group=group))
            if nnz_local:
                ops.append(dist.P2POp(dist.isend, ci.contiguous(), dst, group=group))
                if has_x:
                    ops.append(dist.P2POp(dist.isend, cx.contiguous(), dst, group=group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return (Cp, Ci, Cx) if rank in receivers else None


def sharded_multiply(dA, dB, bounds, rank: int, gather: Optional[str] = "root", group=None, device="cuda",
                     dB_local=None):
    """C(:, J_rank) = A * B(:, J_rank) on this rank (A replicated, no communication), then the
    final gather.

    ``dB_local`` may hold the pre-sliced column block B(:, J_rank) (slice once, multiply many
    times).  ``gather``: None leaves C column-distributed; "root" assembles the whole product on
    rank 0; "all" on every rank.  Returns (local DeviceMatrix, gathered) where gathered is None or
    a tuple of torch tensors (Cp, Ci, Cx).
    """
    import torch
    import torch.distributed as dist
    import csparse_cuda as cc
    j0, j1 = int(bounds[rank]), int(bounds[rank + 1])
    dBl = dB_local if dB_local is not None else dB.col_slice(j0, j1)
    dCl = cc.cs_multiply(dA, dBl)
    if dB_local is None:
        dBl.free()
    if gather is None:
        return dCl, None
    if gather is True:
        gather = "root"
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    p_ptr, i_ptr, x_ptr = dCl.device_pointers()
    # wrap the result's device arrays as torch tensors without copying
    cp = _as_tensor(p_ptr, dCl.n + 1, torch.int32, device)
    ci = _as_tensor(i_ptr, max(dCl.nnz, 1), torch.int32, device)[: dCl.nnz]
    cx = _as_tensor(x_ptr, max(dCl.nnz, 1), torch.float64, device)[: dCl.nnz] if dCl.has_values else None
    return dCl, gather_columns(cp, ci, cx, bounds, rank, world, gather, group, device)


def _as_tensor(ptr: int, count: int, dtype, device):
    """Zero-copy torch view of a device buffer owned by libcsparse_b200 (the handle must outlive it)."""
    import torch

    class _Mem:
        pass
    itemsize = torch.empty(0, dtype=dtype).element_size()
    typestr = {torch.int32: "<i4", torch.float64: "<f8", torch.int64: "<i8"}[dtype]
    holder = _Mem()
    holder.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False),
                                       "version": 2, "strides": None}
    del itemsize
    return torch.as_tensor(holder, device=device)
