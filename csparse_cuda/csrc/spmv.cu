// spmv.cu -- cs_gaxpy (csparse.py:1199-1213): y += A*x, as an atomic-free,
// row-parallel SpMV over the CSR view of A (= the CSC arrays of A', built once by
// cs_transpose and cached on the handle).
//
// The reference walks columns and scatters into y; per output row that is
//     y[i] = (((y[i] + a_1 x_j1) + a_2 x_j2) + ...)      products rounded first,
// in ascending storage order.  Both kernels multiply with __dmul_rn and add with
// __dadd_rn (never FMA).  k_spmv_stream keeps exactly that order per row, so it
// reproduces the reference bit for bit; k_spmv_merge splits long rows between
// threads and is only within the 1e-12 normwise contract.
//
//   k_spmv_stream  short / regular rows: a CTA owns R consecutive rows, streams
//                  their col/val slice with coalesced 128-bit loads, parks the
//                  products in shared memory, then one thread per row sums them
//                  in order.  HBM-bound: 12 B/nnz + 4 B/row + y (16 B/row) + x.
//   k_spmv_merge   power-law rows: merge-path split of (row ends, nonzeros) into
//                  equal tiles per CTA and equal runs per thread, carry-out of
//                  the unfinished row fixed up by k_merge_fixup (deterministic).
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <type_traits>

struct SpmvPlan {
    int kind = 0;            // 1 stream (TMA-staged), 2 merge, 3 stream (plain loads), 4 split (long rows / short rows)
    int rows_per_cta = 0;    // stream
    bool long_rows = false;  // stream: which stage-ring shape (TsLong / TsShort)
    int merge_ctas = 0;      // merge
    int *merge_part = nullptr;      // a-coordinate (row) at the start diagonal of each CTA, merge_ctas+1
    int *carry_row = nullptr;       // per CTA: row of the unfinished tail
    double *carry_val = nullptr;    // per CTA: its partial sum
    csi max_len = 0;
    // host pipeline of csb200_gaxpy: largest column index used by each row chunk (cached)
    int chunk_rows = 0, chunk_count = 0;
    int chunk_maxcol[8] = {0};
    // split: rows binned by length -- long rows cut into items of <= SPLIT_CHUNK entries (one warp
    // each), mid rows (eight lanes each), short rows (one thread each); all read the CSR arrays in place
    int n_items = 0, n_long = 0, n_mid = 0, n_short = 0;
    int4 *items = nullptr;          // {first entry, end, row, unused}
    int *long_list = nullptr;       // n_long: the long rows
    int *long_ptr = nullptr;        // n_long + 1: items of every long row
    double *partial = nullptr;      // n_items
    int *mid_list = nullptr;        // n_mid
    int *short_list = nullptr;      // n_short
};

namespace csb {

constexpr int SP_THREADS = 512;
constexpr int SP_TILE = 4096;          // products parked per CTA (32 KB)

constexpr int MP_THREADS = 256;
#ifndef MP_ITEMS_DEF
#define MP_ITEMS_DEF 7
#endif
#ifndef MP_MINB
#define MP_MINB 5
#endif
constexpr int MP_ITEMS = MP_ITEMS_DEF;                      // odd: the threads' runs of products start 56 B apart, so the
                                                 // 16 lanes of a 64-bit shared-memory phase hit 16 bank pairs
constexpr int MP_TILE = MP_THREADS * MP_ITEMS;   // merge items per CTA

// Products of the slice [base, base+cnt) of (col, val) with x, into prods[0..cnt).
// 128-bit loads on the 4-entry-aligned interior, scalar at the ragged ends.
template <int THREADS>
__device__ __forceinline__ void stream_products(const csi *__restrict__ col, const double *__restrict__ val,
                                                const double *__restrict__ x, int base, int cnt,
                                                double *prods)
{
    const int end = base + cnt;
    const int a0 = base & ~3;                       // aligned start (may precede base)
    for (int k = a0 + threadIdx.x * 4; k < end; k += THREADS * 4) {
        if (k >= base && k + 3 < end) {
            const int4 c = ldg_stream(reinterpret_cast<const int4 *>(col + k));
            const double2 v0 = ldg_stream(reinterpret_cast<const double2 *>(val + k));
            const double2 v1 = ldg_stream(reinterpret_cast<const double2 *>(val + k + 2));
            const double x0 = __ldg(x + c.x), x1 = __ldg(x + c.y), x2 = __ldg(x + c.z), x3 = __ldg(x + c.w);
            double *o = prods + (k - base);
            o[0] = __dmul_rn(v0.x, x0);
            o[1] = __dmul_rn(v0.y, x1);
            o[2] = __dmul_rn(v1.x, x2);
            o[3] = __dmul_rn(v1.y, x3);
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                const int q = k + e;
                if (q >= base && q < end) prods[q - base] = __dmul_rn(val[q], __ldg(x + col[q]));
            }
        }
    }
}

__global__ void __launch_bounds__(SP_THREADS)
k_spmv_stream(int m, const csi *__restrict__ rowptr, const csi *__restrict__ col,
              const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y, int R)
{
    __shared__ double prods[SP_TILE];
    __shared__ int srp[SP_THREADS + 1];
    const int r0 = blockIdx.x * R;
    const int nrows = min(R, m - r0);
    for (int k = threadIdx.x; k <= nrows; k += SP_THREADS) srp[k] = rowptr[r0 + k];
    __syncthreads();
    const int base = srp[0];
    const int cnt = srp[nrows] - base;
    if (cnt <= SP_TILE) {
        stream_products<SP_THREADS>(col, val, x, base, cnt, prods);
        __syncthreads();
        if (threadIdx.x < nrows) {
            const int b = srp[threadIdx.x] - base, e = srp[threadIdx.x + 1] - base;
            if (e > b) {
                double s = y[r0 + threadIdx.x];
                for (int k = b; k < e; k++) s = __dadd_rn(s, prods[k]);
                y[r0 + threadIdx.x] = s;
            }
        }
    } else {
        // irregular block of rows: one warp per row, lanes stride the row
        const int lane = threadIdx.x & 31;
        for (int row = threadIdx.x >> 5; row < nrows; row += SP_THREADS / 32) {
            const int b = srp[row], e = srp[row + 1];
            double s = 0.0;
            for (int k = b + lane; k < e; k += 32) s = __dadd_rn(s, __dmul_rn(val[k], __ldg(x + col[k])));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_down_sync(0xffffffffu, s, o));
            if (lane == 0 && e > b) y[r0 + row] = __dadd_rn(y[r0 + row], s);
        }
    }
}

// ---- TMA-staged persistent row-stream kernel ------------------------------------------
// One CTA per SM slot, looping over row blocks.  A producer warp stages each block's
// rowptr / col / val slices into a ring of shared-memory stages with 1-D bulk copies
// (cp.async.bulk, completion on an mbarrier); 16 consumer warps turn a landed stage
// into products (x gathered through L1/L2), then one thread per row adds its products
// in storage order.  The copy engine keeps several stages in flight per SM, so HBM
// never waits for the dependent x gather or the per-row reduction.
constexpr int TS_CONSUMERS = 512;               // consumer threads == max rows per block
constexpr int TS_THREADS = TS_CONSUMERS + 32;   // + producer warp
constexpr int TS_RP = TS_CONSUMERS + 4;         // staged row pointers (multiple of 4)
// Two shapes of the stage ring, both ~110 KB (two CTAs per SM): short rows (<= 12 entries on
// average) run best with four stages of 2176 nonzeros, longer rows with two stages of 3200
// (measured on lap2d 4096^2: 0.884 vs 0.825 of the copy roofline; st27 128^3: 0.563 vs 0.602).
template <int TS_TILE, int TS_STAGES>
struct TsShape {
    static constexpr int tile = TS_TILE, stages = TS_STAGES;
    static constexpr int stage_bytes = TS_TILE * 12 + TS_RP * 4;
    static constexpr int smem = TS_STAGES * stage_bytes + 64;
};
using TsShort = TsShape<2176, 4>;
using TsLong = TsShape<3200, 2>;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity)
{
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// ---- fused halo exchange (multi-GPU row-block sharding) ----------------
// Each rank owns a slice of x inside a local window [lo halo | own | hi halo].  Instead of an NCCL
// send/recv per step, the ONE persistent SpMV launch pulls the two halos from the neighbours'
// windows over NVLink (peer-mapped pointers): CTA 0 tells the neighbours "my x of this epoch is
// final", waits for theirs, copies the halo lines, raises a local flag and acknowledges the pull.
// Row blocks that read halo entries are moved to the END of the persistent sweep and wait for the
// flag (set long before they are reached); the launch ends only when both neighbours have pulled,
// so the caller may overwrite x as soon as the kernel has completed on its stream.
// Comm block (ints): [0] x ready, set by the lo neighbour  [1] x ready, set by the hi neighbour
//                    [2] pulled, set by the lo neighbour   [3] pulled, set by the hi neighbour
//                    [4] halo landed (local)               [5] a wait timed out (diagnostic)
struct HaloArgs {
    int *comm;
    int *peer_comm[2];
    const double *peer_x[2];
    double *dst[2];
    int cnt[2];
    int epoch;
    int top_blocks, bot_blocks;
};

__device__ __forceinline__ int ld_acquire_sys(const int *p)
{
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(int *p, int v)
{
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p)
{
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}
// spin until *p >= epoch; gives up after ~2 s of SM clocks (a dead neighbour must not hang the GPU)
__device__ __forceinline__ void halo_wait(const int *p, int epoch, int *err)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(p) < epoch) {
        if (clock64() - t0 > 4000000000LL) { *err = 1; break; }
        __nanosleep(64);
    }
}
// executed by `nthreads` threads (ids tid) that can meet on named barrier 1
template <int NTHREADS>
__device__ __forceinline__ void halo_pull(const HaloArgs &h, int tid)
{
    if (tid == 0) {
        if (h.peer_comm[0]) st_release_sys(h.peer_comm[0] + 1, h.epoch);     // I am its hi neighbour
        if (h.peer_comm[1]) st_release_sys(h.peer_comm[1] + 0, h.epoch);     // I am its lo neighbour
        if (h.peer_comm[0]) halo_wait(h.comm + 0, h.epoch, h.comm + 5);
        if (h.peer_comm[1]) halo_wait(h.comm + 1, h.epoch, h.comm + 5);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
#pragma unroll
    for (int side = 0; side < 2; side++)
        if (h.peer_x[side])
        {
            // four remote loads in flight per thread (one NVLink round trip is ~1 us)
            const double *src = h.peer_x[side];
            double *dst = h.dst[side];
            const int cnt = h.cnt[side];
            int k = tid;
            for (; k + 3 * NTHREADS < cnt; k += 4 * NTHREADS) {
                const double v0 = ld_relaxed_sys(src + k), v1 = ld_relaxed_sys(src + k + NTHREADS),
                             v2 = ld_relaxed_sys(src + k + 2 * NTHREADS), v3 = ld_relaxed_sys(src + k + 3 * NTHREADS);
                dst[k] = v0; dst[k + NTHREADS] = v1; dst[k + 2 * NTHREADS] = v2; dst[k + 3 * NTHREADS] = v3;
            }
            for (; k < cnt; k += NTHREADS) dst[k] = ld_relaxed_sys(src + k);
        }
    __threadfence();
    asm volatile("bar.sync 1, %0;" ::"n"(NTHREADS) : "memory");
    if (tid == 0) {
        st_release_sys(h.comm + 4, h.epoch);
        if (h.peer_comm[0]) st_release_sys(h.peer_comm[0] + 3, h.epoch);
        if (h.peer_comm[1]) st_release_sys(h.peer_comm[1] + 2, h.epoch);
    }
}
__device__ __forceinline__ void halo_wait_acks(const HaloArgs &h)
{
    if (h.peer_comm[0]) halo_wait(h.comm + 2, h.epoch, h.comm + 5);
    if (h.peer_comm[1]) halo_wait(h.comm + 3, h.epoch, h.comm + 5);
}

// stand-alone forms for plans without the fused kernel (merge path): pull before, acks after
__global__ void __launch_bounds__(512) k_halo_pull(HaloArgs h) { halo_pull<512>(h, threadIdx.x); }
__global__ void k_halo_acks(HaloArgs h) { halo_wait_acks(h); }

template <class SH, bool HALO>
__global__ void __launch_bounds__(TS_THREADS, 2)
k_spmv_tma(int m, int nblocks, int R, const csi *__restrict__ rowptr, const csi *__restrict__ col,
           const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y, HaloArgs halo)
{
    constexpr int TS_TILE = SH::tile, TS_STAGES = SH::stages, TS_STAGE_BYTES = SH::stage_bytes;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + TS_STAGES * TS_STAGE_BYTES);
    const int tid = threadIdx.x;
    // HALO: CTA 0 (dispatched first; the grid leaves it a resident slot) owns no row blocks.  It runs
    // the halo protocol -- tell the neighbours, wait for them, pull their lines, raise the local
    // flag, wait until they have pulled mine -- so that no worker falls behind its share of the
    // sweep while a neighbour is late.
    const int workers = HALO ? (int)gridDim.x - 1 : (int)gridDim.x;
    const int wix = HALO ? (int)blockIdx.x - 1 : (int)blockIdx.x;      // this CTA among the workers
    if (HALO && blockIdx.x == 0) {
        if (tid < TS_CONSUMERS) {
            halo_pull<TS_CONSUMERS>(halo, tid);
            if (tid == 0) halo_wait_acks(halo);
        }
        return;
    }
    const int niter = (nblocks - wix + workers - 1) / workers;
    // HALO: interior row blocks first, the blocks that read halo entries last
    const int n_int = HALO ? nblocks - halo.top_blocks - halo.bot_blocks : nblocks;
    auto block_of = [&](int q) -> int {
        if (!HALO || q < n_int) return (HALO ? halo.top_blocks : 0) + q;
        const int qb = q - n_int;
        return qb < halo.top_blocks ? qb : n_int + qb;
    };

    if (tid == 0) {
        for (int s = 0; s < TS_STAGES; s++) {
            mbar_init(smem_u32(&bars[s]), 1);                           // full: producer's expect_tx arrival
            mbar_init(smem_u32(&bars[TS_STAGES + s]), TS_CONSUMERS / 32); // empty: one arrival per consumer warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= TS_CONSUMERS) {
        // ---------------- producer warp: lane 0 issues the bulk copies ----------------
        if (tid == TS_CONSUMERS) {
            for (int it = 0; it < niter; it++) {
                const int stage = it % TS_STAGES;
                if (it >= TS_STAGES) mbar_wait(smem_u32(&bars[TS_STAGES + stage]), ((it / TS_STAGES) - 1) & 1);
                const int blk = block_of(wix + it * workers);
                const int r0 = blk * R;
                const int nrows = min(R, m - r0);
                const int start = rowptr[r0], end = rowptr[r0 + nrows];
                const int a0 = start & ~3;
                const int len = ((end + 3) & ~3) - a0;
                const unsigned rp_bytes = (unsigned)(((nrows + 1 + 3) & ~3) * 4);
                const bool staged = len <= TS_TILE;
                unsigned char *st = smem + stage * TS_STAGE_BYTES;
                const unsigned full = smem_u32(&bars[stage]);
                mbar_expect_tx(full, rp_bytes + (staged ? (unsigned)len * 12u : 0u));
                bulk_g2s(smem_u32(st + TS_TILE * 12), rowptr + r0, rp_bytes, full);
                if (staged && len > 0) {
                    bulk_g2s(smem_u32(st), val + a0, (unsigned)len * 8u, full);
                    bulk_g2s(smem_u32(st + TS_TILE * 8), col + a0, (unsigned)len * 4u, full);
                }
            }
        }
        return;
    }

    // ---------------- consumers ----------------
    const int lane = tid & 31;
    // one row block; EDGE (compile time) = the block reads halo entries, which landed during this
    // launch and are therefore read from L2 -- the interior blocks keep the read-only path untouched
    auto run_block = [&](int it, auto edge_tag) {
        constexpr bool edge = decltype(edge_tag)::value;
        const int stage = it % TS_STAGES;
        const int blk = block_of(wix + it * workers);
        const int r0 = blk * R;
        const int nrows = min(R, m - r0);
        double yv = 0.0;
        if (tid < nrows) yv = y[r0 + tid];                   // issued early; used after the products
        unsigned char *st = smem + stage * TS_STAGE_BYTES;
        double *sval = reinterpret_cast<double *>(st);
        const int *scol = reinterpret_cast<const int *>(st + TS_TILE * 8);
        const int *srp = reinterpret_cast<const int *>(st + TS_TILE * 12);
        mbar_wait(smem_u32(&bars[stage]), (it / TS_STAGES) & 1);
        const int start = srp[0], end = srp[nrows];
        const int a0 = start & ~3;
        const int len = ((end + 3) & ~3) - a0;
        if (len <= TS_TILE) {
            for (int k = tid * 4; k < len; k += TS_CONSUMERS * 4) {
                const int4 c = *reinterpret_cast<const int4 *>(scol + k);
                const double2 v0 = *reinterpret_cast<const double2 *>(sval + k);
                const double2 v1 = *reinterpret_cast<const double2 *>(sval + k + 2);
                const int g = a0 + k;                         // global index of the first of four entries
                if (g >= start && g + 3 < end) {
                    double x0, x1, x2, x3;
                    if (edge) { x0 = __ldcg(x + c.x); x1 = __ldcg(x + c.y); x2 = __ldcg(x + c.z); x3 = __ldcg(x + c.w); }   // L2: the halo landed during this launch
                    else      { x0 = __ldg(x + c.x); x1 = __ldg(x + c.y); x2 = __ldg(x + c.z); x3 = __ldg(x + c.w); }
                    double2 o0, o1;
                    o0.x = __dmul_rn(v0.x, x0); o0.y = __dmul_rn(v0.y, x1);
                    o1.x = __dmul_rn(v1.x, x2); o1.y = __dmul_rn(v1.y, x3);
                    *reinterpret_cast<double2 *>(sval + k) = o0;
                    *reinterpret_cast<double2 *>(sval + k + 2) = o1;
                } else {                                      // ragged ends: entries of neighbouring blocks / padding
                    auto ldx = [&](const double *p) { return edge ? __ldcg(p) : __ldg(p); };
                    if (g >= start && g < end) sval[k] = __dmul_rn(v0.x, ldx(x + c.x));
                    if (g + 1 >= start && g + 1 < end) sval[k + 1] = __dmul_rn(v0.y, ldx(x + c.y));
                    if (g + 2 >= start && g + 2 < end) sval[k + 2] = __dmul_rn(v1.x, ldx(x + c.z));
                    if (g + 3 >= start && g + 3 < end) sval[k + 3] = __dmul_rn(v1.y, ldx(x + c.w));
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(TS_CONSUMERS) : "memory");
            if (tid < nrows) {
                const int b = srp[tid] - a0, e = srp[tid + 1] - a0;
                if (e > b) {
                    double s = yv;
                    for (int k = b; k < e; k++) s = __dadd_rn(s, sval[k]);
                    y[r0 + tid] = s;
                }
            }
        } else {
            // block too large for a stage: one warp per row straight from global memory
            for (int row = tid >> 5; row < nrows; row += TS_CONSUMERS / 32) {
                const int b = srp[row], e = srp[row + 1];
                double s = 0.0;
                for (int k = b + lane; k < e; k += 32)
                    s = __dadd_rn(s, __dmul_rn(val[k], edge ? __ldcg(x + col[k]) : __ldg(x + col[k])));
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_down_sync(0xffffffffu, s, o));
                if (lane == 0 && e > b) y[r0 + row] = __dadd_rn(y[r0 + row], s);
            }
        }
        // the stage is about to be overwritten through the async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars[TS_STAGES + stage]));
    };
    int it = 0;
    for (; it < niter && (!HALO || wix + it * workers < n_int); it++) run_block(it, std::false_type{});
    if (HALO && it < niter) {                            // the blocks that read halo entries: they must have landed
        if (tid == 0) halo_wait(halo.comm + 4, halo.epoch, halo.comm + 5);
        asm volatile("bar.sync 1, %0;" ::"n"(TS_CONSUMERS) : "memory");
        for (; it < niter; it++) run_block(it, std::true_type{});
    }
}

static int launch_spmv_tma(bool long_rows, int m, int nblocks, int R, const csi *rowptr, const csi *col,
                           const double *val, const double *x, double *y, cudaStream_t s, const HaloArgs *halo = nullptr)
{
    const int sms = sm_count();
    const int grid = min(nblocks, 2 * sms - (halo ? 1 : 0));   // two resident CTAs per SM (one slot is the halo protocol CTA's)
    HaloArgs h{};
    if (halo) h = *halo;
#define TMA_LAUNCH(SH, HALO)                                                                                  \
    do {                                                                                                      \
        CSB_CUDA(cudaFuncSetAttribute(k_spmv_tma<SH, HALO>, cudaFuncAttributeMaxDynamicSharedMemorySize, SH::smem)); \
        k_spmv_tma<SH, HALO><<<grid + (HALO ? 1 : 0), TS_THREADS, SH::smem, s>>>(m, nblocks, R, rowptr, col, val, x, y, h); \
    } while (0)
    if (long_rows) { if (halo) TMA_LAUNCH(TsLong, true); else TMA_LAUNCH(TsLong, false); }
    else           { if (halo) TMA_LAUNCH(TsShort, true); else TMA_LAUNCH(TsShort, false); }
#undef TMA_LAUNCH
    CSB_LAUNCHED();
    return CSB200_OK;
}

// ---- merge path --------------------------------------------------------------------
// Lists: A = row_end[r] = rowptr[r+1] (r < m), B = 0,1,...,nnz-1.  Item A[a] is
// consumed before B[b] iff row_end[a] <= b ("row a is complete").
__device__ __forceinline__ int merge_search(const int *row_end, int na, int nb, int d, int b_off)
{
    int lo = max(d - nb, 0), hi = min(d, na);
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (row_end[mid] - b_off <= d - 1 - mid) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k_merge_partition(const csi *__restrict__ rowptr, int m, int nnz, int nctas, int *part)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > nctas) return;
    const long long total = (long long)m + nnz;
    const long long d = min((long long)c * MP_TILE, total);
    // coordinates can exceed int only if m+nnz >= 2^31; guarded on the host
    part[c] = merge_search(rowptr + 1, m, nnz, (int)d, 0);
}

__global__ void __launch_bounds__(MP_THREADS, MP_MINB)
k_spmv_merge(int m, int nnz, const csi *__restrict__ rowptr, const csi *__restrict__ col,
             const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y,
             const int *__restrict__ part, int *__restrict__ carry_row, double *__restrict__ carry_val)
{
    // products of the tile's nb nonzeros at the front, row sums of its na finished rows at the
    // back of one buffer: na + nb <= MP_TILE, so they never meet (22 KB per CTA: 8 CTAs per SM)
    __shared__ double prods[MP_TILE];
    __shared__ int re[MP_TILE + 1];          // row ends, local nnz coordinates
    __shared__ int ckey[MP_THREADS];
    __shared__ double cval[MP_THREADS];

    const int c = blockIdx.x;
    const long long total = (long long)m + nnz;
    const int d0 = (int)min((long long)c * MP_TILE, total);
    const int d1 = (int)min((long long)(c + 1) * MP_TILE, total);
    const int a0 = part[c], a1 = part[c + 1];
    const int b0 = d0 - a0, b1 = d1 - a1;
    const int na = a1 - a0;                  // rows completed in this tile
    const int nb = b1 - b0;                  // nonzeros consumed in this tile
    double *rowsum = prods + (MP_TILE - na);

    for (int k = threadIdx.x; k < na; k += MP_THREADS) re[k] = rowptr[a0 + k + 1] - b0;
    stream_products<MP_THREADS>(col, val, x, b0, nb, prods);
    __syncthreads();

    // per-thread run of MP_ITEMS merge items
    const int items = na + nb;
    const int d = min(threadIdx.x * MP_ITEMS, items);
    int a = merge_search(re, na, nb, d, 0);
    int b = d - a;
    double run = 0.0;
#pragma unroll
    for (int it = 0; it < MP_ITEMS; it++) {
        if (a + b < items) {
            if (a < na && re[a] <= b) { rowsum[a] = run; run = 0.0; a++; }
            else                      { run = __dadd_rn(run, prods[b]); b++; }
        }
    }
    // carries: thread t was still accumulating row `a` when its items ran out.  Keys are
    // non-decreasing in t; the partial sums of a run of equal keys are combined by a segmented
    // warp scan plus a walk over the preceding warps' tails -- a fixed order, so the result is
    // deterministic, and no thread ever loops over a whole tile (one row may span all of it).
    {
        const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
        const int key = a;
        double v = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const double vu = __shfl_up_sync(0xffffffffu, v, o);
            const int ku = __shfl_up_sync(0xffffffffu, key, o);
            if (lane >= o && ku == key) v = __dadd_rn(vu, v);
        }
        if (lane == 31) { ckey[wid] = key; cval[wid] = v; }                 // tail of this warp
        if (lane == 0) ckey[MP_THREADS / 32 + wid] = key;                   // head of this warp
        __syncthreads();
        int knext = __shfl_down_sync(0xffffffffu, key, 1);
        if (lane == 31) knext = wid + 1 < MP_THREADS / 32 ? ckey[MP_THREADS / 32 + wid + 1] : -1;
        const int key0 = __shfl_sync(0xffffffffu, key, 0);
        if (knext != key) {                                                  // last thread of its run
            double s = v;
            if (key0 == key)                                                 // run reaches back past lane 0
                for (int w = wid - 1; w >= 0 && ckey[w] == key; w--) s = __dadd_rn(cval[w], s);
            if (key < na) rowsum[key] = __dadd_rn(s, rowsum[key]);           // earlier threads' part comes first
            else { carry_row[c] = a0 + key; carry_val[c] = s; }              // key == na: unfinished tail row
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < na; k += MP_THREADS) {
        const double s = rowsum[k];
        if (s != 0.0) y[a0 + k] = __dadd_rn(y[a0 + k], s);   // empty rows: y untouched, as in the reference
    }
}

// Adds the tile carry-outs.  A row that spans several tiles gets its partial
// sums in tile order from the thread owning the first of them.
__global__ void k_merge_fixup(int nctas, int m, const int *__restrict__ carry_row,
                              const double *__restrict__ carry_val, double *__restrict__ y)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nctas) return;
    const int row = carry_row[c];
    if (row >= m) return;                                    // tile ended exactly at the end
    if (c > 0 && carry_row[c - 1] == row) return;
    double s = carry_val[c];
    for (int u = c + 1; u < nctas && carry_row[u] == row; u++) s = __dadd_rn(s, carry_val[u]);
    if (s != 0.0) y[row] = __dadd_rn(y[row], s);
}

__global__ void k_max_len(const csi *__restrict__ rowptr, int m, int *out)
{
    int mx = 0;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < m; r += gridDim.x * blockDim.x)
        mx = max(mx, rowptr[r + 1] - rowptr[r]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

// ---- split plan: power-law rows ----------------------------------------------------------------
// A row with hundreds of entries needs no shared memory at all: a warp streams it 32 entries at a
// time (coalesced col / val, x gathered through L1), every lane keeps its own running sums and the
// warp reduces once at the end.  Rows are therefore binned by length, once per handle, and every bin
// reads the CSR arrays in place:
//   long   (>= 64 entries: 80+ % of the nonzeros of the R-MAT matrix) cut into items of at
//          most 1024 entries, one warp per item; the items' sums are parked and added to y
//          per row in item order by a small second kernel (deterministic, like the merge path's
//          carry fix-up);
//   mid    (16 .. 63 entries) eight lanes per row, four rows per warp;
//   short  (1 .. 15 entries) one thread per row.
// (cuts swept on R-MAT 2^24: 64 / 8 / 2048 -> 1.227 ms, 64 / 16 / 1024 -> 1.176 ms, 32 / 8 / 2048 -> 1.289 ms)
// Empty rows are in no list: y is untouched there, as in the reference.
struct SplitCuts { int lng, mid, chunk; };      // rows >= lng: long; mid .. lng-1: mid; 1 .. mid-1: short; items of <= chunk entries
static SplitCuts split_cuts()
{
    static const SplitCuts c = {getenv("CSB200_SPLIT_LONG") ? atoi(getenv("CSB200_SPLIT_LONG")) : 64,
                                getenv("CSB200_SPLIT_MID") ? atoi(getenv("CSB200_SPLIT_MID")) : 16,
                                getenv("CSB200_SPLIT_CHUNK") ? atoi(getenv("CSB200_SPLIT_CHUNK")) : 1024};
    return c;
}

__global__ void k_split_count(int m, const csi *__restrict__ rowptr, SplitCuts cut, int *__restrict__ nitems,
                              int *__restrict__ islong, int *__restrict__ ismid, int *__restrict__ isshort)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int len = rowptr[r + 1] - rowptr[r];
    const bool lg = len >= cut.lng;
    nitems[r] = lg ? (len + cut.chunk - 1) / cut.chunk : 0;
    islong[r] = lg ? 1 : 0;
    ismid[r] = (!lg && len >= cut.mid) ? 1 : 0;
    isshort[r] = (len > 0 && len < cut.mid) ? 1 : 0;
}

// iptr / lptr / mptr / sptr: exclusive scans of the four arrays above
__global__ void k_split_fill(int m, const csi *__restrict__ rowptr, SplitCuts cut, const int *__restrict__ iptr,
                             const int *__restrict__ lptr, const int *__restrict__ mptr, const int *__restrict__ sptr,
                             int4 *__restrict__ items, int *__restrict__ long_list, int *__restrict__ long_ptr,
                             int *__restrict__ mid_list, int *__restrict__ short_list)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int b = rowptr[r], e = rowptr[r + 1];
    const int len = e - b;
    if (len >= cut.lng) {
        const int li = lptr[r];
        long_list[li] = r;
        long_ptr[li] = iptr[r];
        int k = iptr[r];
        // equal items: a row of 2049 entries becomes 1025 + 1024, not 2048 + 1
        const int n = (len + cut.chunk - 1) / cut.chunk;
        for (int c = 0; c < n; c++, k++) {
            const int ib = b + (int)((long long)len * c / n), ie = b + (int)((long long)len * (c + 1) / n);
            items[k] = make_int4(ib, ie, r, 0);
        }
    } else if (len >= cut.mid) {
        mid_list[mptr[r]] = r;
    } else if (len > 0) {
        short_list[sptr[r]] = r;
    }
}

template <int UNR>
__global__ void __launch_bounds__(256)
k_spmv_long(int n_items, const int4 *__restrict__ items, const csi *__restrict__ col, const double *__restrict__ val,
            const double *__restrict__ x, double *__restrict__ partial)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * 8;
    for (int it = blockIdx.x * 8 + (threadIdx.x >> 5); it < n_items; it += nwarps) {
        const int4 im = items[it];
        double acc[UNR];
#pragma unroll
        for (int u = 0; u < UNR; u++) acc[u] = 0.0;
        // UNR independent gathers in flight per lane all the way: the last round is predicated instead
        // of falling back to one entry at a time (the kernel is latency bound: long scoreboard 90 per issue)
        for (int p = im.x + lane; p < im.y; p += 32 * UNR) {
            int c[UNR];
            double v[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                const bool ok = p + 32 * u < im.y;
                c[u] = ok ? ldg_stream(col + p + 32 * u) : -1;
                v[u] = ok ? ldg_stream(val + p + 32 * u) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < UNR; u++)
                if (c[u] >= 0) acc[u] = __dadd_rn(acc[u], __dmul_rn(v[u], __ldg(x + c[u])));
        }
        double s = acc[0];
#pragma unroll
        for (int u = 1; u < UNR; u++) s = __dadd_rn(s, acc[u]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
        if (lane == 0) partial[it] = s;
    }
}

__global__ void k_long_fix(int n_long, const int *__restrict__ long_rows, const int *__restrict__ long_ptr,
                           const double *__restrict__ partial, double *__restrict__ y)
{
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= n_long) return;
    double s = 0.0;
    for (int k = long_ptr[li]; k < long_ptr[li + 1]; k++) s = __dadd_rn(s, partial[k]);
    const int r = long_rows[li];
    y[r] = __dadd_rn(y[r], s);
}

// eight lanes per row
__global__ void __launch_bounds__(256)
k_spmv_mid(int n_mid, const int *__restrict__ mid_list, const csi *__restrict__ rowptr, const csi *__restrict__ col,
           const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
    const int gi = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 3);
    const int sub = threadIdx.x & 7;
    double s = 0.0;
    int r = -1;
    if (gi < n_mid) {
        r = mid_list[gi];
        const int e = rowptr[r + 1];
        double s2 = 0.0;
        for (int p = rowptr[r] + sub; p < e; p += 32) {     // four predicated gathers in flight per lane
            int c[4];
            double v[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const bool ok = p + 8 * u < e;
                c[u] = ok ? col[p + 8 * u] : -1;
                v[u] = ok ? val[p + 8 * u] : 0.0;
            }
            if (c[0] >= 0) s = __dadd_rn(s, __dmul_rn(v[0], __ldg(x + c[0])));
            if (c[1] >= 0) s2 = __dadd_rn(s2, __dmul_rn(v[1], __ldg(x + c[1])));
            if (c[2] >= 0) s = __dadd_rn(s, __dmul_rn(v[2], __ldg(x + c[2])));
            if (c[3] >= 0) s2 = __dadd_rn(s2, __dmul_rn(v[3], __ldg(x + c[3])));
        }
        s = __dadd_rn(s, s2);
    }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, o));
    if (r >= 0 && sub == 0) y[r] = __dadd_rn(y[r], s);
}

// one thread per row, entries in storage order
__global__ void __launch_bounds__(256)
k_spmv_short(int n_short, const int *__restrict__ short_list, const csi *__restrict__ rowptr, const csi *__restrict__ col,
             const double *__restrict__ val, const double *__restrict__ x, double *__restrict__ y)
{
    const int gi = blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n_short) return;
    const int r = short_list[gi];
    const int e = rowptr[r + 1];
    double s = y[r];
    for (int p = rowptr[r]; p < e; p++) s = __dadd_rn(s, __dmul_rn(val[p], __ldg(x + col[p])));   // the reference's order (csparse.py:1210-1212)
    y[r] = s;
}

// side stream of the split plan (per thread and device, like the host pipeline's copy streams)
struct SplitStreams {
    cudaStream_t side = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr;
    int device = -1;
    int init()
    {
        int dev = 0;
        CSB_CUDA(cudaGetDevice(&dev));
        if (device == dev) return CSB200_OK;
        CSB_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));     // (highest stream priority for this one: 1.23 against 1.14 ms)
        CSB_CUDA(cudaEventCreateWithFlags(&fork, cudaEventDisableTiming));
        CSB_CUDA(cudaEventCreateWithFlags(&join, cudaEventDisableTiming));
        device = dev;
        return CSB200_OK;
    }
};
static SplitStreams &split_streams()
{
    static thread_local SplitStreams ss;
    return ss;
}

static int build_split(csb200_mat *AT, SpmvPlan *pl)
{
    const int m = AT->n;
    cudaStream_t s = stream();
    DevBuf<int> nitems, islong, ismid, isshort, iptr, lptr, mptr, sptr;
    DevBuf<long long> tot;
    const size_t cap = (size_t)m + 1;
    CSB_TRY(nitems.alloc(cap)); CSB_TRY(islong.alloc(cap)); CSB_TRY(ismid.alloc(cap)); CSB_TRY(isshort.alloc(cap));
    CSB_TRY(iptr.alloc(cap)); CSB_TRY(lptr.alloc(cap)); CSB_TRY(mptr.alloc(cap)); CSB_TRY(sptr.alloc(cap));
    CSB_TRY(tot.alloc(4));
    const SplitCuts cut = split_cuts();
    k_split_count<<<ceil_div(m, 256), 256, 0, s>>>(m, AT->p, cut, nitems.ptr, islong.ptr, ismid.ptr, isshort.ptr);
    CSB_LAUNCHED();
    CSB_TRY(launch_excl_scan(iptr.ptr, nitems.ptr, m, tot.ptr, nullptr));
    CSB_TRY(launch_excl_scan(lptr.ptr, islong.ptr, m, tot.ptr + 1, nullptr));
    CSB_TRY(launch_excl_scan(mptr.ptr, ismid.ptr, m, tot.ptr + 2, nullptr));
    CSB_TRY(launch_excl_scan(sptr.ptr, isshort.ptr, m, tot.ptr + 3, nullptr));
    long long h[4] = {0, 0, 0, 0};
    CSB_CUDA(cudaMemcpyAsync(h, tot.ptr, sizeof(h), cudaMemcpyDeviceToHost, s));
    CSB_CUDA(cudaStreamSynchronize(s));
    pl->n_items = (int)h[0];
    pl->n_long = (int)h[1];
    pl->n_mid = (int)h[2];
    pl->n_short = (int)h[3];
    CSB_TRY(dev_alloc(&pl->items, (size_t)pl->n_items + 1));
    CSB_TRY(dev_alloc(&pl->long_list, (size_t)pl->n_long + 1));
    CSB_TRY(dev_alloc(&pl->long_ptr, (size_t)pl->n_long + 1));
    CSB_TRY(dev_alloc(&pl->partial, (size_t)pl->n_items + 1));
    CSB_TRY(dev_alloc(&pl->mid_list, (size_t)pl->n_mid + 1));
    CSB_TRY(dev_alloc(&pl->short_list, (size_t)pl->n_short + 1));
    k_split_fill<<<ceil_div(m, 256), 256, 0, s>>>(m, AT->p, cut, iptr.ptr, lptr.ptr, mptr.ptr, sptr.ptr, pl->items,
                                                  pl->long_list, pl->long_ptr, pl->mid_list, pl->short_list);
    CSB_LAUNCHED();
    CSB_CUDA(cudaMemcpyAsync(pl->long_ptr + pl->n_long, &pl->n_items, sizeof(int), cudaMemcpyHostToDevice, s));
    CSB_CUDA(cudaStreamSynchronize(s));      // n_items is a host variable of the plan: copied before it can change
    return CSB200_OK;
}

void spmv_plan_free(SpmvPlan *pl)
{
    if (!pl) return;
    dev_free(pl->merge_part);
    dev_free(pl->carry_row);
    dev_free(pl->carry_val);
    dev_free(pl->items);
    dev_free(pl->long_list);
    dev_free(pl->long_ptr);
    dev_free(pl->partial);
    dev_free(pl->mid_list);
    dev_free(pl->short_list);
    delete pl;
}

// AT: CSC of A' == CSR of A.  rows = AT->n, cols = AT->m.
int spmv_build_plan(csb200_mat *AT)
{
    if (AT->plan && (AT->forced_plan == 0 || AT->plan->kind == AT->forced_plan)) return CSB200_OK;
    if (AT->plan) { spmv_plan_free(AT->plan); AT->plan = nullptr; }
    const int m = AT->n;
    const long long nnz = AT->nnz;
    // longest row decides the kernel
    int h_max = 0;
    if (m > 0) {
        DevBuf<int> d_max;
        CSB_TRY(d_max.alloc(1));
        CSB_CUDA(cudaMemsetAsync(d_max.ptr, 0, sizeof(int), stream()));
        k_max_len<<<min(ceil_div(m, 256), sm_count() * 8), 256, 0, stream()>>>(AT->p, m, d_max.ptr);
        CSB_LAUNCHED();
        CSB_CUDA(cudaMemcpyAsync(&h_max, d_max.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream()));
        CSB_CUDA(cudaStreamSynchronize(stream()));
    }
    SpmvPlan *pl = new SpmvPlan();
    pl->max_len = h_max;
    const double avg = m > 0 ? (double)nnz / m : 0.0;
    int kind = (h_max <= 64 || h_max <= 4.0 * avg + 16.0) ? 1 : 2;
    if (kind == 2 && nnz >= (1 << 20)) kind = 4;                    // power-law rows: binned by length (1.23 ms against the merge path's 1.87 on R-MAT 2^24)
    if (kind == 2 && (long long)m + nnz >= 0x7fffffffLL - MP_TILE) kind = 1;     // merge coordinates are int32
    if (AT->forced_plan) kind = AT->forced_plan;
    pl->kind = kind;
    if (kind == 1) {
        // rows per block: a multiple of 4 (16-byte aligned rowptr slices), sized so that an
        // average block fills ~90 % of a stage
        pl->long_rows = avg > 12.0;
        int R = (int)(0.9 * (pl->long_rows ? TsLong::tile : TsShort::tile) / (avg > 1.0 ? avg : 1.0));
        R = R > TS_CONSUMERS ? TS_CONSUMERS : R;
        R &= ~3;
        pl->rows_per_cta = R < 4 ? 4 : R;
    } else if (kind == 3) {
        int R = (int)(0.8 * SP_TILE / (avg > 1.0 ? avg : 1.0));
        pl->rows_per_cta = R < 1 ? 1 : (R > SP_THREADS ? SP_THREADS : R);
    } else if (kind == 4) {
        int st = build_split(AT, pl);
        if (st != CSB200_OK) { spmv_plan_free(pl); return st; }
    } else {
        const long long total = (long long)m + nnz;
        pl->merge_ctas = (int)((total + MP_TILE - 1) / MP_TILE);
        int st = dev_alloc(&pl->merge_part, (size_t)pl->merge_ctas + 1);
        if (st == CSB200_OK) st = dev_alloc(&pl->carry_row, (size_t)pl->merge_ctas + 1);
        if (st == CSB200_OK) st = dev_alloc(&pl->carry_val, (size_t)pl->merge_ctas + 1);
        if (st != CSB200_OK) { spmv_plan_free(pl); return st; }
        k_merge_partition<<<ceil_div(pl->merge_ctas + 1, 256), 256, 0, stream()>>>(
            AT->p, m, (int)nnz, pl->merge_ctas, pl->merge_part);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { spmv_plan_free(pl); return set_error(CSB200_ERR_CUDA, "k_merge_partition: %s", cudaGetErrorString(e)); }
    }
    AT->plan = pl;
    return CSB200_OK;
}

// y[0..AT->n) += AT' * x[0..AT->m)
int spmv_run(csb200_mat *AT, const double *d_x, double *d_y)
{
    if (!AT->x) return set_error(CSB200_ERR_ARG, "cs_gaxpy: matrix has no values");
    CSB_TRY(spmv_build_plan(AT));
    const int m = AT->n;
    if (m == 0 || AT->nnz == 0) return CSB200_OK;
    SpmvPlan *pl = AT->plan;
    if (pl->kind == 1) {
        const int R = pl->rows_per_cta;
        const int nblocks = ceil_div(m, R);
        CSB_TRY(launch_spmv_tma(pl->long_rows, m, nblocks, R, AT->p, AT->i, AT->x, d_x, d_y, stream()));
    } else if (pl->kind == 3) {
        const int R = pl->rows_per_cta;
        k_spmv_stream<<<ceil_div(m, R), SP_THREADS, 0, stream()>>>(m, AT->p, AT->i, AT->x, d_x, d_y, R);
        CSB_LAUNCHED();
    } else if (pl->kind == 4) {
        // the three bins write disjoint rows of y: the mid and short kernels run beside the long one
        // on a second stream (forked and joined with events), so their tails overlap
        cudaStream_t s = stream();
        SplitStreams &ss = split_streams();
        CSB_TRY(ss.init());
        CSB_CUDA(cudaEventRecord(ss.fork, s));
        CSB_CUDA(cudaStreamWaitEvent(ss.side, ss.fork, 0));
        if (pl->n_items > 0) {
            // four predicated gathers in flight per lane, twelve CTAs per SM (4 or 8 in flight x 4..16 CTAs:
            // 1.136 .. 1.146 ms on R-MAT 2^24 -- flat; an L2 evict-first policy on the streamed entries: no change)
            const int grid = min(ceil_div(pl->n_items, 8), sm_count() * 12);
            k_spmv_long<4><<<grid, 256, 0, s>>>(pl->n_items, pl->items, AT->i, AT->x, d_x, pl->partial);
            CSB_LAUNCHED();
            k_long_fix<<<ceil_div(pl->n_long, 256), 256, 0, s>>>(pl->n_long, pl->long_list, pl->long_ptr, pl->partial, d_y);
            CSB_LAUNCHED();
        }
        if (pl->n_mid > 0) {
            k_spmv_mid<<<ceil_div((long long)pl->n_mid * 8, 256), 256, 0, ss.side>>>(pl->n_mid, pl->mid_list, AT->p, AT->i, AT->x, d_x, d_y);
            CSB_LAUNCHED();
        }
        if (pl->n_short > 0) {
            k_spmv_short<<<ceil_div(pl->n_short, 256), 256, 0, ss.side>>>(pl->n_short, pl->short_list, AT->p, AT->i, AT->x, d_x, d_y);
            CSB_LAUNCHED();
        }
        CSB_CUDA(cudaEventRecord(ss.join, ss.side));
        CSB_CUDA(cudaStreamWaitEvent(s, ss.join, 0));
    } else {
        k_spmv_merge<<<pl->merge_ctas, MP_THREADS, 0, stream()>>>(m, (int)AT->nnz, AT->p, AT->i, AT->x, d_x, d_y,
                                                                   pl->merge_part, pl->carry_row, pl->carry_val);
        CSB_LAUNCHED();
        k_merge_fixup<<<ceil_div(pl->merge_ctas, 256), 256, 0, stream()>>>(pl->merge_ctas, m, pl->carry_row,
                                                                            pl->carry_val, d_y);
        CSB_LAUNCHED();
    }
    return CSB200_OK;
}

// Rows [ra, rb) only, for the chunked host pipeline of csb200_gaxpy: possible when the plan is
// the row-stream kernel (row blocks are independent) and ra is a multiple of the plan's block
// height.  *align receives that height.
int spmv_rows_align(csb200_mat *AT, int *align)
{
    CSB_TRY(spmv_build_plan(AT));
    *align = AT->plan->kind == 1 ? AT->plan->rows_per_cta : 0;
    return CSB200_OK;
}

__global__ void k_chunk_maxcol(const csi *__restrict__ rowptr, const csi *__restrict__ col, int m, int rows,
                               int *__restrict__ out)
{
    // one CTA per (chunk, slice): the entries of chunk c are col[rowptr[c*rows] .. rowptr[min(m,(c+1)*rows)])
    const int c = blockIdx.y;
    const int ra = min(m, c * rows), rb = min(m, ra + rows);
    const long long b = rowptr[ra], e = rowptr[rb];
    int mx = -1;
    for (long long k = b + (long long)blockIdx.x * blockDim.x + threadIdx.x; k < e; k += (long long)gridDim.x * blockDim.x)
        mx = max(mx, col[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0 && mx >= 0) atomicMax(&out[c], mx);
}

// largest column each chunk of `rows` rows touches (host array of `count` ints, cached in the plan)
int spmv_chunk_maxcol(csb200_mat *AT, int rows, int count, const int **out)
{
    SpmvPlan *pl = AT->plan;
    if (!pl || count > 8) return set_error(CSB200_ERR_ARG, "spmv_chunk_maxcol: bad arguments");
    if (pl->chunk_rows != rows || pl->chunk_count != count) {
        DevBuf<int> d;
        CSB_TRY(d.alloc(8));
        CSB_CUDA(cudaMemsetAsync(d.ptr, 0xff, 8 * sizeof(int), stream()));
        k_chunk_maxcol<<<dim3(sm_count(), count), 256, 0, stream()>>>(AT->p, AT->i, AT->n, rows, d.ptr);
        CSB_LAUNCHED();
        CSB_CUDA(cudaMemcpyAsync(pl->chunk_maxcol, d.ptr, 8 * sizeof(int), cudaMemcpyDeviceToHost, stream()));
        CSB_CUDA(cudaStreamSynchronize(stream()));
        pl->chunk_rows = rows;
        pl->chunk_count = count;
    }
    *out = pl->chunk_maxcol;
    return CSB200_OK;
}

int spmv_run_rows(csb200_mat *AT, const double *d_x, double *d_y, int ra, int rb, cudaStream_t s)
{
    SpmvPlan *pl = AT->plan;
    if (!pl || pl->kind != 1 || ra % pl->rows_per_cta != 0) return set_error(CSB200_ERR_ARG, "spmv_run_rows: bad range");
    if (rb <= ra) return CSB200_OK;
    const int R = pl->rows_per_cta;
    const int nblocks = ceil_div(rb - ra, R);
    return launch_spmv_tma(pl->long_rows, rb - ra, nblocks, R, AT->p + ra, AT->i, AT->x, d_x, d_y + ra, s);
}

}  // namespace csb

namespace csb {
int spmv_plan_kind(const SpmvPlan *pl) { return pl ? pl->kind : 0; }
}

// ---- the halo object: x window + comm block of one rank, peer-mapped views of its neighbours' ----
struct csb200_halo {
    double *window = nullptr;       // cudaMalloc (IPC-exportable), `count` doubles
    int *comm = nullptr;            // 8 ints
    long long count = 0;
    int device = 0;
    int epoch = 0;
    // per side (0 = lo neighbour, 1 = hi neighbour)
    void *peer_window_base[2] = {nullptr, nullptr};   // mapping to close (IPC) or null (same-process peer)
    void *peer_comm_base[2] = {nullptr, nullptr};
    const double *peer_x[2] = {nullptr, nullptr};
    int *peer_comm[2] = {nullptr, nullptr};
    long long local_first[2] = {0, 0};
    int cnt[2] = {0, 0};
};

using namespace csb;

static HaloArgs halo_args(csb200_halo *h, int top_blocks, int bot_blocks)
{
    HaloArgs a{};
    a.comm = h->comm;
    for (int s = 0; s < 2; s++) {
        a.peer_comm[s] = h->peer_comm[s];
        a.peer_x[s] = h->peer_x[s];
        a.dst[s] = h->window + h->local_first[s];
        a.cnt[s] = h->cnt[s];
    }
    a.epoch = h->epoch;
    a.top_blocks = top_blocks;
    a.bot_blocks = bot_blocks;
    return a;
}

namespace csb {
// the stand-alone halo protocol for the host-buffer pipeline of csb200_gaxpy_halo (api.cu): pull the
// neighbours' lines into the window now / wait until the neighbours have pulled mine
int halo_pull_launch(csb200_halo *h, cudaStream_t s)
{
    h->epoch++;
    const HaloArgs a = halo_args(h, 0, 0);
    k_halo_pull<<<1, 512, 0, s>>>(a);
    CSB_LAUNCHED();
    return CSB200_OK;
}
int halo_acks_launch(csb200_halo *h, cudaStream_t s)
{
    const HaloArgs a = halo_args(h, 0, 0);
    k_halo_acks<<<1, 1, 0, s>>>(a);
    CSB_LAUNCHED();
    return CSB200_OK;
}
double *halo_window_ptr(csb200_halo *h) { return h->window; }
long long halo_window_count(csb200_halo *h) { return h->count; }
}  // namespace csb

extern "C" {

int csb200_halo_create(int64_t count, csb200_halo **out)
{
    if (!out || count < 0) return set_error(CSB200_ERR_ARG, "halo_create: bad arguments");
    *out = nullptr;
    CSB_TRY(ensure_device());
    csb200_halo *h = new csb200_halo();
    h->count = count;
    cudaGetDevice(&h->device);
    cudaError_t e = cudaMalloc((void **)&h->window, (size_t)(count > 0 ? count : 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void **)&h->comm, 8 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemset(h->window, 0, (size_t)(count > 0 ? count : 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemset(h->comm, 0, 8 * sizeof(int));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaFree(h->window); cudaFree(h->comm);
        delete h;
        return set_error(e == cudaErrorMemoryAllocation ? CSB200_ERR_NOMEM : CSB200_ERR_CUDA, "halo_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return CSB200_OK;
}

int csb200_halo_window(csb200_halo *h, double **d_window)
{
    if (!h || !d_window) return set_error(CSB200_ERR_ARG, "halo_window: null argument");
    *d_window = h->window;
    return CSB200_OK;
}

int csb200_halo_export(csb200_halo *h, void *handles)
{
    if (!h || !handles) return set_error(CSB200_ERR_ARG, "halo_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "two 64-byte handles");
    cudaIpcMemHandle_t hw, hc;
    CSB_CUDA(cudaIpcGetMemHandle(&hw, h->window));
    CSB_CUDA(cudaIpcGetMemHandle(&hc, h->comm));
    memcpy(handles, &hw, 64);
    memcpy(static_cast<char *>(handles) + 64, &hc, 64);
    return CSB200_OK;
}

static int halo_set_side(csb200_halo *h, int side, const double *pw, int *pc, int64_t peer_first, int64_t count, int64_t local_first)
{
    if (count < 0 || count > 0x7fffffff || local_first < 0 || local_first + count > h->count || peer_first < 0)
        return set_error(CSB200_ERR_ARG, "halo_connect: range outside the window");
    h->peer_x[side] = pw + peer_first;
    h->peer_comm[side] = pc;
    h->local_first[side] = local_first;
    h->cnt[side] = (int)count;
    return CSB200_OK;
}

int csb200_halo_connect(csb200_halo *h, int side, const void *peer_handles, int64_t peer_first, int64_t count,
                        int64_t local_first)
{
    if (!h || !peer_handles || side < 0 || side > 1) return set_error(CSB200_ERR_ARG, "halo_connect: bad arguments");
    cudaIpcMemHandle_t hw, hc;
    memcpy(&hw, peer_handles, 64);
    memcpy(&hc, static_cast<const char *>(peer_handles) + 64, 64);
    void *pw = nullptr, *pc = nullptr;
    CSB_CUDA(cudaIpcOpenMemHandle(&pw, hw, cudaIpcMemLazyEnablePeerAccess));
    cudaError_t e = cudaIpcOpenMemHandle(&pc, hc, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaIpcCloseMemHandle(pw);
        return set_error(CSB200_ERR_CUDA, "halo_connect: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
    }
    h->peer_window_base[side] = pw;
    h->peer_comm_base[side] = pc;
    return halo_set_side(h, side, static_cast<const double *>(pw), static_cast<int *>(pc), peer_first, count, local_first);
}

int csb200_halo_connect_local(csb200_halo *h, int side, csb200_halo *peer, int64_t peer_first, int64_t count,
                              int64_t local_first)
{
    if (!h || !peer || side < 0 || side > 1) return set_error(CSB200_ERR_ARG, "halo_connect_local: bad arguments");
    if (peer->device != h->device) {
        int can = 0;
        CSB_CUDA(cudaDeviceCanAccessPeer(&can, h->device, peer->device));
        if (!can) return set_error(CSB200_ERR_CUDA, "halo_connect_local: no peer access between devices %d and %d", h->device, peer->device);
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(h->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(peer->device, 0);
        cudaSetDevice(cur);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            return set_error(CSB200_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    return halo_set_side(h, side, peer->window, peer->comm, peer_first, count, local_first);
}

// y[0..AT.n) += AT' * window, the halos pulled from the neighbours inside the same launch.
// top_rows / bot_rows: rows at the two ends of the block that read halo entries.
int csb200_gaxpy_halo_dev(csb200_mat *AT, csb200_halo *h, double *d_y, csi top_rows, csi bot_rows)
{
    if (!AT || !h || !d_y || top_rows < 0 || bot_rows < 0) return set_error(CSB200_ERR_ARG, "cs_gaxpy: null argument");
    if (!AT->x) return set_error(CSB200_ERR_ARG, "cs_gaxpy: matrix has no values");
    if (AT->m > h->count) return set_error(CSB200_ERR_ARG, "cs_gaxpy: the x window is shorter than the block's column range");
    ArenaScope arena_scope;
    CSB_TRY(spmv_build_plan(AT));
    const int m = AT->n;
    h->epoch++;
    SpmvPlan *pl = AT->plan;
    cudaStream_t s = stream();
    if (m > 0 && AT->nnz > 0 && pl->kind == 1) {
        const int R = pl->rows_per_cta;
        const int nblocks = ceil_div(m, R);
        // blocks that hold a row reading halo entries: the first ceil(top_rows / R), and every block
        // from the one holding row m - bot_rows on (the last block may be short)
        int tb = ceil_div(top_rows, R);
        int bb = bot_rows > 0 ? nblocks - max(0, m - (int)bot_rows) / R : 0;
        if (tb + bb > nblocks) { tb = nblocks; bb = 0; }
        const HaloArgs a = halo_args(h, tb, bb);
        return launch_spmv_tma(pl->long_rows, m, nblocks, R, AT->p, AT->i, AT->x, h->window, d_y, s, &a);
    }
    const HaloArgs a = halo_args(h, 0, 0);
    k_halo_pull<<<1, 512, 0, s>>>(a);
    CSB_LAUNCHED();
    CSB_TRY(spmv_run(AT, h->window, d_y));
    k_halo_acks<<<1, 1, 0, s>>>(a);
    CSB_LAUNCHED();
    return CSB200_OK;
}

int csb200_halo_status(csb200_halo *h, int *timed_out)
{
    if (!h || !timed_out) return set_error(CSB200_ERR_ARG, "halo_status: null argument");
    int v = 0;
    CSB_CUDA(cudaMemcpyAsync(&v, h->comm + 5, sizeof(int), cudaMemcpyDeviceToHost, stream()));
    CSB_CUDA(cudaStreamSynchronize(stream()));
    *timed_out = v;
    return CSB200_OK;
}

int csb200_halo_free(csb200_halo *h)
{
    if (!h) return CSB200_OK;
    cudaDeviceSynchronize();
    for (int s = 0; s < 2; s++) {
        if (h->peer_window_base[s]) cudaIpcCloseMemHandle(h->peer_window_base[s]);
        if (h->peer_comm_base[s]) cudaIpcCloseMemHandle(h->peer_comm_base[s]);
    }
    cudaFree(h->window);
    cudaFree(h->comm);
    cudaGetLastError();
    delete h;
    return CSB200_OK;
}

}  // extern "C"
