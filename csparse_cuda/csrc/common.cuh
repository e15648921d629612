// common.cuh -- shared host/device helpers of libcsparse_b200.so (sm_100a only)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/csparse_b200.h"

namespace csb {

// ---- per-thread library state ---------------------------------------------
struct ThreadState {
    cudaStream_t stream = nullptr;   // legacy default stream unless csb200_set_stream
    char err[640] = {0};
    int64_t last_flops = 0;
    // path switches of the tests / benchmarks (csb200_*_force_path): per thread, like the stream
    int force_transpose = 0;          // 1: always the radix sort, 2: automatic choice without the mirror path, 3: 2 in L2-sized slabs, 4: 2 fused into one persistent launch
    int multiply_ordered = 0;         // 1: always the reference's discovery order
    int multiply_blocked_version = 0; // 2 / 3: which blocked numeric kernel (0 = default)
    int multiply_templates = 0;       // pattern-class path: 0 automatic, 1 off, 2 on at any size, 3 like 2 without the lane-per-column kernel
    int64_t last_templated = 0;       // columns the last cs_multiply of this thread formed from class templates
    int add_force_spgemm = 0;         // 1: cs_add on the SpGEMM kernels even for canonical operands
    long long *host_scalar = nullptr; // 64 bytes of pinned, device-visible host memory: a kernel's scalar result lands here without a copy
};
ThreadState &tls();
extern std::atomic<int64_t> g_launches;

int set_error(int status, const char *fmt, ...);

#define CSB_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return csb::set_error(CSB200_ERR_CUDA, "%s: %s (%s:%d)", #expr,              \
                                  cudaGetErrorString(e__), __FILE__, __LINE__);          \
    } while (0)

#define CSB_TRY(expr)                                                                    \
    do {                                                                                 \
        int s__ = (expr);                                                                \
        if (s__ != CSB200_OK) return s__;                                                \
    } while (0)

// call right after a <<<>>> launch
#define CSB_LAUNCHED()                                                                   \
    do {                                                                                 \
        csb::g_launches.fetch_add(1, std::memory_order_relaxed);                         \
        CSB_CUDA(cudaGetLastError());                                                    \
    } while (0)

inline cudaStream_t stream() { return tls().stream; }

// SMs of the current device (queried once per device; grids are sized in multiples of it)
int sm_count();
#ifdef __CUDACC__
// grid of a persistent kernel: every CTA resident at once (SMs x CTAs that fit one SM)
template <class K>
inline int resident_grid(K kern, int threads, size_t smem = 0)
{
    int per = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, threads, smem) != cudaSuccess) { cudaGetLastError(); per = 1; }
    return sm_count() * (per > 0 ? per : 1);
}
#endif

// stream-ordered allocation from the device's default memory pool (cached: the
// release threshold is raised once per device in ensure_device()).  Results (the p / i / x of a
// handle) and per-handle caches come from here.
int ensure_device();
template <class T>
inline int dev_alloc(T **p, size_t count)
{
    *p = nullptr;
    if (count == 0) count = 1;
    CSB_TRY(ensure_device());
    cudaError_t e = cudaMallocAsync((void **)p, count * sizeof(T), stream());
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(e == cudaErrorMemoryAllocation ? CSB200_ERR_NOMEM : CSB200_ERR_CUDA,
                         "cudaMallocAsync(%zu bytes): %s", count * sizeof(T), cudaGetErrorString(e));
    }
    return CSB200_OK;
}
inline void dev_free(void *p)
{
    if (p) cudaFreeAsync(p, stream());
}

// ---- per-thread workspace arena ------------------------------------------------------------
// Temporaries of an API call (histograms, bucket intermediates, work lists, block tables) are
// bump-allocated from ONE block per thread that only ever grows: a call never touches the
// stream-ordered pool for them, so the pool sees nothing but same-sized results and the first
// call after an upload costs what the tenth does.  Stream order makes reuse safe: the next call's
// kernels queue behind this call's on the same stream (a thread that switches streams pays one
// synchronize).  Requests the block cannot hold fall back to the pool for that call and raise
// the high-water mark, so the next outermost scope grows the block once.
struct Arena {
    char *base = nullptr;
    size_t cap = 0, off = 0;
    size_t voff = 0;          // like off, but also counting the requests that overflowed to the pool
    size_t want = 0;          // high-water mark of voff: what the block should hold
    int depth = 0;
    int device = -1;
    cudaStream_t last_stream = nullptr;
};
Arena &arena();
int arena_enter();            // ArenaScope's constructor: grows the block at the outermost level if needed
void arena_hint(size_t bytes);   // "this call will need about this much": grows right away when nothing is in use
struct ArenaScope {
    size_t saved_off, saved_voff;
    ArenaScope() { Arena &a = arena(); arena_enter(); saved_off = a.off; saved_voff = a.voff; a.depth++; }
    ~ArenaScope()
    {
        Arena &a = arena();
        a.depth--;
        a.off = saved_off;
        a.voff = saved_voff;
    }
};

// RAII for temporaries inside an API call: from the arena when a scope is open and the block has
// room, else from the pool (freed stream-ordered on destruction)
template <class T>
struct DevBuf {
    T *ptr = nullptr;
    bool pooled = false;
    ~DevBuf() { if (pooled) dev_free(ptr); }
    int alloc(size_t count)
    {
        if (count == 0) count = 1;
        const size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        Arena &a = arena();
        if (a.depth > 0) {
            a.voff += bytes;
            if (a.voff > a.want) a.want = a.voff;
            if (a.base && a.off + bytes <= a.cap) {
                ptr = reinterpret_cast<T *>(a.base + a.off);
                a.off += bytes;
                return CSB200_OK;
            }
        }
        pooled = true;
        return dev_alloc(&ptr, count);
    }
    operator T *() const { return ptr; }
};

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Every p / i / x array of a handle is over-allocated by MAT_PAD elements so that
// 16-byte-granular bulk (TMA) copies may read a few elements past the logical end.
constexpr size_t MAT_PAD = 8;

// ---- kernels exported between translation units ----------------------------
// scan.cu : p[0..n] = exclusive scan(c), c[i] <- p[i]; d_total (int64) and
//           d_max (max element, may be null) are device scalars.
int launch_excl_scan(csi *d_p, csi *d_c, csi n, long long *d_total, int *d_max);
// radix.cu : stable LSD radix sort of (key, a, v) triples by key in [0, nkeys) and the
//            column pointers Cp[0..nkeys] of the result; a == nullptr derives the int payload
//            from Ap as "the column holding this position"; v == nullptr sorts pattern only.
int stable_sort_by_key(long long nnz, int nkeys, const int *key, const int *a, const csi *Ap, int ncols,
                       const double *v, csi *Cp, csi *a_out, double *v_out);

}  // namespace csb

// ---- the opaque matrix handle ------------------------------------------------
struct SpmvPlan;
struct csb200_mat {
    csi m = 0, n = 0;
    int64_t nnz = 0;
    csi *p = nullptr;      // n+1, device
    csi *i = nullptr;      // max(nnz,1), device
    double *x = nullptr;   // max(nnz,1) or null (pattern only)
    int device = 0;
    // lazily computed facts / caches (handles are logically immutable)
    int canon = -1;              // 1: every column strictly increasing (=> no duplicate entries)
    csi max_col_len = -1;        // longest column
    int mirror = -1;             // 1: square, canonical and pattern-symmetric (cs_transpose's one-pass path); 0: known not to be
    int wide_rows = -1;          // 1: a 4096-entry tile's rows span more buckets than the bucket sort's window (cs_transpose takes the radix sort)
    csb200_mat *csr = nullptr;   // cached transpose with values == CSR view of this matrix
    SpmvPlan *plan = nullptr;    // gaxpy plan over this matrix interpreted as a CSR view (columns = rows)
    int forced_plan = 0;
    // cs_multiply's symbolic phase: every column as (32-row block, bit mask) pairs, stored at the
    // column's own offset p[j] with c32_len[j] pairs (spgemm.cu, built on first use as left factor)
    csi *c32_blk = nullptr;
    unsigned *c32_mask = nullptr;
    csi *c32_len = nullptr;
    // cs_multiply's pattern classes (spgemm_tpl.cuh): cls[k] = class of column k's pattern relative to
    // k, or -1; cls_state -1 not computed, 0 too many classes (unstructured), 1 usable
    csi *cls = nullptr;
    int cls_state = -1;
    int cls_count = 0;
    // entry-major copy of the values, soa_x[e * n + k] = e-th value of column k (k_num_soa): soa_state
    // -1 not built, 0 does not qualify, 1 usable with soa_len slots per column
    double *soa_x = nullptr;
    int soa_state = -1;
    int soa_len = 0;
};

// ---- device helpers -----------------------------------------------------------
#ifdef __CUDACC__
namespace csb {

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ unsigned lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }

// Lanes holding the same NBITS-bit value as this lane (valid lanes only), from NBITS + 1 ballots.
// The hardware MATCH.ANY takes ~900 cycles per warp at full occupancy on sm_100 (it was a third of the
// radix pass's stall samples, profiles/r2_notes.md); ballots issue at full rate and do not depend on
// the shared-memory chain of the ranking loop, so consecutive rounds overlap.
template <int NBITS>
__device__ __forceinline__ unsigned match_bits(int d, bool valid)
{
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < NBITS; b++) {
        const bool bit = (d & (1 << b)) != 0;                 // LOP3 into a predicate, VOTE, SEL, LOP3: four instructions per bit
        const unsigned bal = __ballot_sync(0xffffffffu, bit);
        peers &= bal ^ (bit ? 0u : 0xffffffffu);
    }
    return peers;
}

// streaming (read-once) loads: bypass L1 allocation
__device__ __forceinline__ int4 ldg_stream(const int4 *p)
{
    int4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ int ldg_stream(const int *p)
{
    int r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ double ldg_stream(const double *p)
{
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}

// 256-bit accesses (sm_100: LDG/STG.E.ENL2.256): four consecutive doubles of one thread as ONE
// request, so that a warp whose lanes own 32 contiguous bytes each moves whole 32-byte sectors per
// instruction (two 16-byte accesses with a 32-byte lane stride touch every sector twice, half-filled).
// p must be 32-byte aligned.
__device__ __forceinline__ void ldg_stream4(const double *p, double &a, double &b, double &c, double &d)
{
    asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}
__device__ __forceinline__ void stg4(double *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// largest j in [lo, hi] with a[j] <= v  (a non-decreasing, a[lo] <= v assumed)
__device__ __forceinline__ int upper_row(const int *a, int lo, int hi, int v)
{
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (a[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

}  // namespace csb
#endif
