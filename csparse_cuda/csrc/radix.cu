// radix.cu -- stable LSD radix sort of (key, int payload, double payload) triples by an
// integer key in [0, nkeys), plus the column-pointer array of the result.
//
// This is the reference's stable counting-sort scatter (cs_transpose, csparse.py:2305-2314;
// cs_compress, csparse.py:663-671) for inputs whose key distribution defeats the two-level
// bucket sort of transpose.cu (power-law rows) or whose entries are in no particular order
// (triplets).  Stability gives the reference's order inside every output column -- ascending
// source position -- by construction, so nothing has to be repaired afterwards.
//
//   k_rs_hist    digit histograms of every pass in one read of the keys
//   k_rs_starts  256-entry exclusive scans -> first output slot of every digit
//   k_rs_pass    one 8-bit pass, "onesweep" style: a tile of 4096 entries is ranked with
//                __match_any (each warp owns 256 consecutive entries and a private counter
//                per digit, so ranks follow the source order without atomics), the tile's
//                digit counts are chained to the previous tiles with a decoupled look-back
//                (one thread per digit), then every entry goes straight to its final slot
//                of this pass.  Intermediate passes move 16-byte {key, a, v} records.
//   k_rs_bounds  Cp[r] = first slot whose key is >= r (binary search in the sorted keys)
//
// Algorithmic bytes per entry: 4 (histogram) + 32 per pass (read + write a record) + 8
// (sorted keys written and searched); ceil(log2(nkeys) / 8) passes.
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace csb {

#ifndef RS_THREADS_DEF
#define RS_THREADS_DEF 512
#endif
constexpr int RS_THREADS = RS_THREADS_DEF;
constexpr int RS_WARPS = RS_THREADS / 32;
#ifndef RS_EPT_DEF
#define RS_EPT_DEF 16
#endif
constexpr int RS_EPT = RS_EPT_DEF;              // entries per thread
constexpr int RS_TILE = RS_THREADS * RS_EPT;    // 4096 entries per tile
constexpr int RS_SEG = 32 * RS_EPT;             // consecutive entries owned by one warp
constexpr int RS_BINS = 256;
constexpr int RS_MAX_PASSES = 4;

struct __align__(16) RsRec { int key; int a; double v; };
struct __align__(8) RsRecP { int key; int a; };

constexpr unsigned long long RS_AGG = 1ull << 62, RS_PREFIX = 2ull << 62;

// One 128-bit (64-bit for pattern-only records) access per record: a plain struct copy of
// {int, int, double} compiles to two 64-bit accesses.
__device__ __forceinline__ RsRec ld_rec(const RsRec *p)
{
    const int4 t = __ldg(reinterpret_cast<const int4 *>(p));
    RsRec r;
    r.key = t.x; r.a = t.y; r.v = __hiloint2double(t.w, t.z);
    return r;
}
__device__ __forceinline__ RsRecP ld_rec(const RsRecP *p)
{
    const int2 t = __ldg(reinterpret_cast<const int2 *>(p));
    RsRecP r;
    r.key = t.x; r.a = t.y;
    return r;
}
__device__ __forceinline__ void st_rec(RsRec *p, const RsRec &r)
{
    *reinterpret_cast<int4 *>(p) = make_int4(r.key, r.a, __double2loint(r.v), __double2hiint(r.v));
}
__device__ __forceinline__ void st_rec(RsRecP *p, const RsRecP &r)
{
    *reinterpret_cast<int2 *>(p) = make_int2(r.key, r.a);
}

__global__ void __launch_bounds__(256)
k_rs_hist(const int *__restrict__ key, long long nnz, int npasses, unsigned long long *__restrict__ hist)
{
    __shared__ unsigned h[RS_MAX_PASSES][RS_BINS];
    for (int k = threadIdx.x; k < RS_MAX_PASSES * RS_BINS; k += 256) (&h[0][0])[k] = 0;
    __syncthreads();
    const long long stride = (long long)gridDim.x * 256 * 4;
    for (long long p = ((long long)blockIdx.x * 256 + threadIdx.x) * 4; p < nnz; p += stride) {
        int k4[4];
        if (p + 3 < nnz) {
            const int4 r = ldg_stream(reinterpret_cast<const int4 *>(key + p));
            k4[0] = r.x; k4[1] = r.y; k4[2] = r.z; k4[3] = r.w;
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) k4[e] = p + e < nnz ? key[p + e] : -1;
        }
#pragma unroll
        for (int e = 0; e < 4; e++)
            if (k4[e] >= 0)
                for (int q = 0; q < npasses; q++) atomicAdd(&h[q][(k4[e] >> (8 * q)) & 255], 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < npasses * RS_BINS; k += 256) {
        const unsigned c = (&h[0][0])[k];
        if (c) atomicAdd(&hist[k], (unsigned long long)c);
    }
}

__global__ void __launch_bounds__(RS_BINS)
k_rs_starts(int npasses, unsigned long long *hist)
{
    __shared__ unsigned long long wsum[RS_BINS / 32];
    const int d = threadIdx.x, lane = d & 31, wid = d >> 5;
    for (int q = 0; q < npasses; q++) {
        const unsigned long long c = hist[q * RS_BINS + d];
        unsigned long long inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (lane == 31) wsum[wid] = inc;
        __syncthreads();
        unsigned long long before = inc - c;
        for (int w = 0; w < wid; w++) before += wsum[w];
        hist[q * RS_BINS + d] = before;
        __syncthreads();
    }
}

// Decoupled look-back of one digit: the sum of the counts of the tiles before `tile`.  A walk that
// reads one predecessor per step is a chain of dependent L2 round trips (the whole CTA waits at the
// next barrier meanwhile); `lb` statuses are read at once and consumed in order, so a walk of k tiles
// costs ceil(k / lb) round trips.  A status that is not published yet ends the batch: the walk resumes there.
// the same walk, one predecessor per step, starting from a status word that is already on its way
__device__ __forceinline__ long long rs_look_back_from(volatile unsigned long long *status, unsigned tile, int d,
                                                       unsigned long long w)
{
    long long prefix = 0;
    for (long long t = (long long)tile - 1;; ) {
        volatile unsigned long long *pst = status + (size_t)t * RS_BINS + d;
        while ((w >> 62) == 0) w = *pst;
        prefix += (long long)(w & 0xffffffffull);
        if ((w & (2ull << 62)) || --t < 0) break;
        w = status[(size_t)t * RS_BINS + d];
    }
    return prefix;
}
constexpr int RS_LB_MAX = 8;
__device__ __forceinline__ long long rs_look_back(volatile unsigned long long *status, unsigned tile, int d, int lb)
{
    long long prefix = 0;
    long long t = (long long)tile - 1;
    if (lb <= 1) {                                        // one predecessor per step
        for (; t >= 0; t--) {
            volatile unsigned long long *pst = status + (size_t)t * RS_BINS + d;
            unsigned long long w = *pst;
            while ((w >> 62) == 0) w = *pst;
            prefix += (long long)(w & 0xffffffffull);
            if (w & (2ull << 62)) break;
        }
        return prefix;
    }
    while (t >= 0) {
        unsigned long long w[RS_LB_MAX];
#pragma unroll
        for (int u = 0; u < RS_LB_MAX; u++)
            w[u] = (u < lb && t - u >= 0) ? status[(size_t)(t - u) * RS_BINS + d] : (2ull << 62);    // past the first tile: an empty prefix
        int used = 0;
        bool done = false;
#pragma unroll
        for (int u = 0; u < RS_LB_MAX; u++) {
            if (done || u != used || u >= lb) continue;
            if ((w[u] >> 62) == 0) continue;              // not published yet: stop consuming here
            prefix += (long long)(w[u] & 0xffffffffull);
            used = u + 1;
            if (w[u] & RS_PREFIX) done = true;
        }
        if (done) break;
        t -= used;
    }
    return prefix;
}

// SRC: 0 = separate key / a / v arrays, 1 = records, 2 = separate key / v arrays with the
//      int payload derived as "the column of Ap that holds this position" (cs_transpose)
// DST: 0 = separate arrays (last pass; the sorted keys are written too), 1 = records
// dynamic shared memory of k_rs_pass: the per-warp counters, then either one word per entry (records:
// digit and source slot) or the tile's keys in sorted order plus a 16-bit source slot (separate arrays:
// the key then needs no gather of its own)
constexpr int rs_pass_smem(int src) { return RS_WARPS * RS_BINS * (int)sizeof(int) + RS_TILE * (src == 1 ? 4 : 6); }

#ifndef RS_MINB_DEF
#define RS_MINB_DEF (1024 / RS_THREADS)
#endif
template <int SRC, int DST, bool VALUES>
__global__ void __launch_bounds__(RS_THREADS, RS_MINB_DEF)
k_rs_pass(long long nnz, int shift,
          const int *__restrict__ key_in, const int *__restrict__ a_in, const double *__restrict__ v_in,
          const void *__restrict__ rec_in, const csi *__restrict__ Ap, int ncols,
          int *__restrict__ key_out, int *__restrict__ a_out, double *__restrict__ v_out, void *__restrict__ rec_out,
          const unsigned long long *__restrict__ digit_start, volatile unsigned long long *status, unsigned *ticket, int lb)
{
    using Rec = typename std::conditional<VALUES, RsRec, RsRecP>::type;
    extern __shared__ __align__(16) unsigned char rs_dyn[];      // rs_pass_smem() bytes: the counters and the permutation
    int (*cnt)[RS_BINS] = reinterpret_cast<int (*)[RS_BINS]>(rs_dyn);
    unsigned *perm = reinterpret_cast<unsigned *>(rs_dyn + RS_WARPS * RS_BINS * sizeof(int));   // SRC == 1: (digit << 16) | source slot inside the tile; else the key
    unsigned short *perm16 = reinterpret_cast<unsigned short *>(perm + RS_TILE);                   // SRC != 1: the source slot
    __shared__ long long gbase[RS_BINS];            // first global slot of the digit's run of this tile MINUS its first slot in the tile
    __shared__ int toff[RS_BINS];
    __shared__ int wtot[RS_BINS / 32];
    __shared__ unsigned s_tile;
    __shared__ int s_col[2];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = lanemask_lt();
    if (tid == 0) s_tile = atomicAdd(ticket, 1u);
    for (int k = tid; k < RS_WARPS * RS_BINS; k += RS_THREADS) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const unsigned tile = s_tile;
    const long long base = (long long)tile * RS_TILE;
    const long long seg = base + wid * RS_SEG;
    const Rec *rin = reinterpret_cast<const Rec *>(rec_in);
    if (SRC == 2 && tid < 2) {
        // columns holding the first and the last position of this tile
        const long long pe = tid == 0 ? base : min(nnz, base + RS_TILE) - 1;
        s_col[tid] = upper_row(Ap, 0, ncols, (int)pe);
    }

    // ---- ranks in source order --------------------------------------------------------
    int key[RS_EPT];
    int rank[RS_EPT];
#pragma unroll
    for (int u = 0; u < RS_EPT; u++) {
        const long long e = seg + u * 32 + lane;
        key[u] = -1;
        if (e < nnz) key[u] = SRC == 1 ? rin[e].key : key_in[e];
    }
#pragma unroll
    for (int u = 0; u < RS_EPT; u++) {
        // warp-uniform control flow: lanes past the end are in nobody's peer mask (their own is unused)
        const bool valid = key[u] >= 0;
        const int d = (key[u] >> shift) & (RS_BINS - 1);
        const unsigned peers = match_bits<8>(d, valid);
        const int leader = __ffs(peers) - 1;
        int r = 0;
        if (valid && lane == leader) { r = cnt[wid][d]; cnt[wid][d] = r + __popc(peers); }
        rank[u] = __shfl_sync(0xffffffffu, r, leader) + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();

    // ---- tile counts -> published; the look-back itself waits until the permutation is written --------
    // The first predecessor's status is requested right after this tile's counts are out and consumed
    // only after the permutation phase: one L2 round trip of the walk is hidden, and the predecessors
    // have had that much longer to publish their prefixes.
    int my_total = 0, before = 0;
    unsigned long long w_first = 0;
    long long early_prefix = 0;
    if (tid < RS_BINS) {
        const int d = tid;
        int sum = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) { const int c = cnt[w][d]; cnt[w][d] = sum; sum += c; }
        my_total = sum;
        volatile unsigned long long *mine = status + (size_t)tile * RS_BINS + d;
        if (tile == 0) {
            *mine = RS_PREFIX | (unsigned long long)sum;
        } else {
            *mine = RS_AGG | (unsigned long long)sum;
            if (lb == 0) {                         // A/B: the whole walk right here
                early_prefix = rs_look_back(status, tile, d, 1);
                *mine = RS_PREFIX | (unsigned long long)(early_prefix + sum);
            } else {
                w_first = status[(size_t)(tile - 1) * RS_BINS + d];
            }
        }
    }
    // first slot of every digit inside the tile's own sorted order (exclusive scan of the totals)
    {
        int inc = my_total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (tid < RS_BINS && lane == 31) wtot[wid] = inc;
        __syncthreads();
        if (tid < RS_BINS) {
            before = inc - my_total;
            for (int w = 0; w < wid; w++) before += wtot[w];
            toff[tid] = before;
        }
    }
    __syncthreads();

    // ---- the tile's permutation goes through shared memory so that consecutive threads write
    //      consecutive slots of a digit's run (full sectors / bursts instead of 16-byte shards)
#pragma unroll
    for (int u = 0; u < RS_EPT; u++) {
        if (key[u] < 0) continue;
        const int d = (key[u] >> shift) & (RS_BINS - 1);
        const int tpos = toff[d] + cnt[wid][d] + rank[u];
        if (SRC == 1) {
            perm[tpos] = ((unsigned)d << 16) | (unsigned)(wid * RS_SEG + u * 32 + lane);
        } else {
            perm[tpos] = (unsigned)key[u];
            perm16[tpos] = (unsigned short)(wid * RS_SEG + u * 32 + lane);
        }
    }
    // ---- look-back -> global base of every digit ----------------------------------------------------------
    if (tid < RS_BINS) {
        const int d = tid;
        long long prefix = 0;
        if (tile > 0 && lb == 0) prefix = early_prefix;
        else if (tile > 0) {
            prefix = lb > 1 ? rs_look_back(status, tile, d, lb) : rs_look_back_from(status, tile, d, w_first);
            status[(size_t)tile * RS_BINS + d] = RS_PREFIX | (unsigned long long)(prefix + my_total);
        }
        gbase[d] = (long long)digit_start[d] + prefix - before;      // slot = gbase[d] + position in the tile
    }
    __syncthreads();
    const int tile_n = (int)min((long long)RS_TILE, nnz - base);
    const int j_lo = SRC == 2 ? s_col[0] : 0, j_hi = SRC == 2 ? s_col[1] : 0;
    // SRC == 2: the column of every entry of the tile, without a binary search per entry -- every
    // non-empty column that starts inside the tile marks its first position, then a running maximum
    // carries the marks forward.  The table (16-bit offsets from the tile's first column) takes the
    // place of the per-warp digit counters, which are no longer needed; a tile spanning more than
    // 65535 columns (long runs of empty columns) keeps the binary search.
    unsigned short *colof = reinterpret_cast<unsigned short *>(&cnt[0][0]);
    static_assert(RS_WARPS * RS_BINS * sizeof(int) >= RS_TILE * sizeof(unsigned short), "the column table fits the counters' space");
    const bool marked = SRC == 2 && j_hi - j_lo < 65535;
    if (marked) {
#pragma unroll
        for (int k = 0; k < RS_EPT; k++) colof[k * RS_THREADS + tid] = 0;
        __syncthreads();
        for (int j = j_lo + 1 + tid; j <= j_hi; j += RS_THREADS) {
            const int a0 = Ap[j];
            const long long q = (long long)a0 - base;
            if (q >= 0 && q < tile_n && Ap[j + 1] > a0) colof[(int)q] = (unsigned short)(j - j_lo);   // one non-empty column per position
        }
        __syncthreads();
        // inclusive running maximum over the tile: RS_EPT consecutive entries per thread
        int v[RS_EPT];
        int run = 0;
#pragma unroll
        for (int k = 0; k < RS_EPT; k++) { run = max(run, (int)colof[tid * RS_EPT + k]); v[k] = run; }
        int inc = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, t); }
        __shared__ int wmax[RS_WARPS];
        if (lane == 31) wmax[wid] = inc;
        __syncthreads();
        int before = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) before = 0;
        for (int w = 0; w < wid; w++) before = max(before, wmax[w]);
#pragma unroll
        for (int k = 0; k < RS_EPT; k++) colof[tid * RS_EPT + k] = (unsigned short)max(v[k], before);
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < RS_EPT; k++) {
        const int tpos = k * RS_THREADS + tid;
        if (tpos >= tile_n) continue;
        const unsigned pw = perm[tpos];
        const int src = SRC == 1 ? (int)(pw & 0xffffu) : (int)perm16[tpos];
        const long long e = base + src;
        const long long pos = gbase[SRC == 1 ? (pw >> 16) : ((pw >> shift) & (RS_BINS - 1))] + tpos;
        int kk, a;
        double v = 0.0;
        if (SRC == 1) {
            const Rec r = ld_rec(rin + e);
            kk = r.key; a = r.a;
            if constexpr (VALUES) v = r.v;
        } else {
            kk = (int)pw;
            a = SRC == 2 ? (marked ? j_lo + (int)colof[src] : upper_row(Ap, j_lo, j_hi, (int)e)) : a_in[e];
            if (VALUES) v = v_in[e];
        }
        if (DST == 1) {
            Rec r;
            r.key = kk; r.a = a;
            if constexpr (VALUES) r.v = v;
            st_rec(reinterpret_cast<Rec *>(rec_out) + pos, r);
        } else {
            key_out[pos] = kk;
            a_out[pos] = a;
            if (VALUES) v_out[pos] = v;
        }
    }
}

// ---- the same pass with the payload routed through shared memory -------------------------------------
// k_rs_pass above routes only the tile's PERMUTATION through shared memory and gathers every record
// from global memory in destination order: one 32-byte sector per entry through L1 (two for the first
// pass, whose key and value live in separate arrays) -- the kernel ran at 80 % of the L1 data stage and
// 45 % of the HBM rate.  Here a thread keeps the records it loaded (coalesced, once) in registers,
// parks them in shared memory at their position in the tile's sorted order, and the tile leaves as
// consecutive 16-byte records of every digit run.  Tile = 4096 entries (64 KB of records), two CTAs per SM.
#ifndef RS2_THREADS_DEF
#define RS2_THREADS_DEF 256
#endif
constexpr int RS2_THREADS = RS2_THREADS_DEF;        // 256: four CTAs per SM, every thread owns a digit during the look-back
constexpr int RS2_WARPS = RS2_THREADS / 32;
constexpr int RS2_EPT = 8;
constexpr int RS2_TILE = RS2_THREADS * RS2_EPT;
constexpr int RS2_SEG = 32 * RS2_EPT;
template <bool VALUES>
constexpr int rs2_smem() { return RS2_WARPS * RS_BINS * 4 + RS_BINS * 8 + RS_BINS * 4 + 64 + RS2_TILE * (VALUES ? 16 : 8); }

template <int SRC, int DST, bool VALUES>
__global__ void __launch_bounds__(RS2_THREADS, 1024 / RS2_THREADS)
k_rs_pass_s(long long nnz, int shift,
            const int *__restrict__ key_in, const int *__restrict__ a_in, const double *__restrict__ v_in,
            const void *__restrict__ rec_in, const csi *__restrict__ Ap, int ncols,
            int *__restrict__ key_out, int *__restrict__ a_out, double *__restrict__ v_out, void *__restrict__ rec_out,
            const unsigned long long *__restrict__ digit_start, volatile unsigned long long *status, unsigned *ticket, int lb)
{
    using Rec = typename std::conditional<VALUES, RsRec, RsRecP>::type;
    extern __shared__ __align__(16) unsigned char rs_smem[];
    int (*cnt)[RS_BINS] = reinterpret_cast<int (*)[RS_BINS]>(rs_smem);
    long long *gbase = reinterpret_cast<long long *>(rs_smem + RS2_WARPS * RS_BINS * 4);
    int *toff = reinterpret_cast<int *>(rs_smem + RS2_WARPS * RS_BINS * 4 + RS_BINS * 8);
    int *misc = toff + RS_BINS;                         // [0..7] wtot, [8] tile, [9..10] first / last column, 16 ints in all
    Rec *srec = reinterpret_cast<Rec *>(rs_smem + RS2_WARPS * RS_BINS * 4 + RS_BINS * 8 + RS_BINS * 4 + 64);
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = lanemask_lt();
    if (tid == 0) misc[8] = (int)atomicAdd(ticket, 1u);
    for (int k = tid; k < RS2_WARPS * RS_BINS; k += RS2_THREADS) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const unsigned tile = (unsigned)misc[8];
    const long long base = (long long)tile * RS2_TILE;
    const long long seg = base + wid * RS2_SEG;
    const int tile_n = (int)min((long long)RS2_TILE, nnz - base);
    const Rec *rin = reinterpret_cast<const Rec *>(rec_in);
    if (SRC == 2 && tid < 2) {
        const long long pe = tid == 0 ? base : base + tile_n - 1;     // columns holding the tile's first / last position
        misc[9 + tid] = upper_row(Ap, 0, ncols, (int)pe);
    }

    // ---- the thread's records, loaded once, coalesced ---------------------------------------------------
    int key[RS2_EPT], pay[RS2_EPT];
    double val[RS2_EPT];
#pragma unroll
    for (int u = 0; u < RS2_EPT; u++) {
        const long long e = seg + u * 32 + lane;
        key[u] = -1; pay[u] = 0; val[u] = 0.0;
        if (e < nnz) {
            if (SRC == 1) {
                const Rec r = ld_rec(rin + e);
                key[u] = r.key; pay[u] = r.a;
                if constexpr (VALUES) val[u] = r.v;
            } else {
                key[u] = key_in[e];
                if (SRC == 0) pay[u] = a_in[e];
                if (VALUES) val[u] = v_in[e];
            }
        }
    }
    // ---- ranks in source order (as in k_rs_pass) ----------------------------------------------------------
    int rank[RS2_EPT];
#pragma unroll
    for (int u = 0; u < RS2_EPT; u++) {
        const bool valid = key[u] >= 0;
        const int d = (key[u] >> shift) & (RS_BINS - 1);
        const unsigned peers = match_bits<8>(d, valid);
        const int leader = __ffs(peers) - 1;
        int r = 0;
        if (valid && lane == leader) { r = cnt[wid][d]; cnt[wid][d] = r + __popc(peers); }
        rank[u] = __shfl_sync(0xffffffffu, r, leader) + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // ---- tile counts -> look-back -> global base of every digit ------------------------------------------------
    int my_total = 0;
    if (tid < RS_BINS) {
        const int d = tid;
        int sum = 0;
#pragma unroll
        for (int w = 0; w < RS2_WARPS; w++) { const int c = cnt[w][d]; cnt[w][d] = sum; sum += c; }
        my_total = sum;
        volatile unsigned long long *mine = status + (size_t)tile * RS_BINS + d;
        long long prefix = 0;
        if (tile == 0) {
            *mine = RS_PREFIX | (unsigned long long)sum;
        } else {
            *mine = RS_AGG | (unsigned long long)sum;
            prefix = rs_look_back(status, tile, d, lb);
            *mine = RS_PREFIX | (unsigned long long)(prefix + sum);
        }
        gbase[d] = (long long)digit_start[d] + prefix;
    }
    {
        int inc = my_total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        if (tid < RS_BINS && lane == 31) misc[wid] = inc;
        __syncthreads();
        if (tid < RS_BINS) {
            int before = inc - my_total;
            for (int w = 0; w < wid; w++) before += misc[w];
            toff[tid] = before;
        }
    }
    __syncthreads();
    // ---- position of every record inside the tile's sorted order (the counters are free afterwards) ------------
    int tpos[RS2_EPT];
#pragma unroll
    for (int u = 0; u < RS2_EPT; u++) {
        const int d = (key[u] >> shift) & (RS_BINS - 1);
        tpos[u] = key[u] >= 0 ? toff[d] + cnt[wid][d] + rank[u] : 0;
    }
    __syncthreads();
    if (SRC == 2) {
        // the column of every entry: marks of the non-empty columns that start inside the tile, carried
        // forward by a running maximum (16-bit offsets from the tile's first column, in the counters' space)
        const int j_lo = misc[9], j_hi = misc[10];
        unsigned short *colof = reinterpret_cast<unsigned short *>(&cnt[0][0]);
        static_assert(RS2_WARPS * RS_BINS * sizeof(int) >= RS2_TILE * sizeof(unsigned short), "the column table fits the counters' space");
        if (j_hi - j_lo < 65535) {
#pragma unroll
            for (int k = 0; k < RS2_EPT; k++) colof[k * RS2_THREADS + tid] = 0;
            __syncthreads();
            for (int j = j_lo + 1 + tid; j <= j_hi; j += RS2_THREADS) {
                const int a0 = Ap[j];
                const long long q = (long long)a0 - base;
                if (q >= 0 && q < tile_n && Ap[j + 1] > a0) colof[(int)q] = (unsigned short)(j - j_lo);
            }
            __syncthreads();
            int v[RS2_EPT];
            int run = 0;
#pragma unroll
            for (int k = 0; k < RS2_EPT; k++) { run = max(run, (int)colof[tid * RS2_EPT + k]); v[k] = run; }
            int inc = run;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc = max(inc, t); }
            if (lane == 31) misc[wid] = inc;             // 16 warps: misc[0..15]... only [0..7] are free: see below
            __syncthreads();
            int before = __shfl_up_sync(0xffffffffu, inc, 1);
            if (lane == 0) before = 0;
            for (int w = 0; w < wid; w++) before = max(before, misc[w]);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < RS2_EPT; k++) colof[tid * RS2_EPT + k] = (unsigned short)max(v[k], before);
            __syncthreads();
#pragma unroll
            for (int u = 0; u < RS2_EPT; u++) pay[u] = j_lo + (int)colof[wid * RS2_SEG + u * 32 + lane];
        } else {
#pragma unroll
            for (int u = 0; u < RS2_EPT; u++)
                if (key[u] >= 0) pay[u] = upper_row(Ap, j_lo, j_hi, (int)(seg + u * 32 + lane));
        }
    }
    // ---- park the records at their sorted positions, then leave in order ---------------------------------------
#pragma unroll
    for (int u = 0; u < RS2_EPT; u++) {
        if (key[u] < 0) continue;
        Rec r;
        r.key = key[u]; r.a = pay[u];
        if constexpr (VALUES) r.v = val[u];
        st_rec(srec + tpos[u], r);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS2_EPT; k++) {
        const int tp = k * RS2_THREADS + tid;
        if (tp >= tile_n) continue;
        const Rec r = srec[tp];
        const int d = (r.key >> shift) & (RS_BINS - 1);
        const long long pos = gbase[d] + (tp - toff[d]);
        if (DST == 1) {
            st_rec(reinterpret_cast<Rec *>(rec_out) + pos, r);
        } else {
            key_out[pos] = r.key;
            a_out[pos] = r.a;
            if constexpr (VALUES) v_out[pos] = r.v;
        }
    }
}

// Cp[r] = number of entries with key < r, r = 0..nkeys (keys sorted ascending)
// (one streamed pass over the keys with the lanes of a warp filling the rows between two different
// neighbours measured the same 0.4 ms on R-MAT 2^24 and lost on matrices without empty rows)
__global__ void k_rs_bounds(const int *__restrict__ keys, long long nnz, int nkeys, csi *__restrict__ Cp)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > nkeys) return;
    long long lo = 0, hi = nnz;            // first index with keys[idx] >= r
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < r) lo = mid + 1; else hi = mid;
    }
    Cp[r] = (csi)lo;
}

#define RS_CUDA(expr) CSB_CUDA(expr)

template <bool VALUES>
static int sort_impl(long long nnz, int nkeys, const int *key, const int *a, const csi *Ap, int ncols,
                     const double *v, csi *Cp, csi *a_out, double *v_out)
{
    cudaStream_t s = stream();
    int bits = 1;
    while (bits < 31 && (1LL << bits) < (long long)nkeys) bits++;
    const int npasses = (bits + 7) / 8;
    static const bool staged = getenv("CSB200_RS_STAGED") && atoi(getenv("CSB200_RS_STAGED"));   // A/B switch: measured slower (12.8 vs 11.3 ms on R-MAT 2^24), off by default
    static const int lb_env = getenv("CSB200_RS_LB") ? atoi(getenv("CSB200_RS_LB")) : 1;    // batched look-back: A/B switch (helps the staged pass only)
    const int lb = lb_env < 0 ? 0 : (lb_env > RS_LB_MAX ? RS_LB_MAX : lb_env);
    const int ntiles = ceil_div(nnz, staged ? RS2_TILE : RS_TILE);
    const size_t recsz = VALUES ? sizeof(RsRec) : sizeof(RsRecP);

    arena_hint((size_t)nnz * 4 + (npasses >= 3 ? 2 : npasses >= 2 ? 1 : 0) * (size_t)nnz * recsz +
               (size_t)ntiles * RS_BINS * 8 + (1 << 16));
    DevBuf<unsigned long long> hist, status;
    DevBuf<unsigned> ticket;
    DevBuf<unsigned char> bufA, bufB;
    DevBuf<int> keys_sorted;
    CSB_TRY(hist.alloc(RS_MAX_PASSES * RS_BINS));
    CSB_TRY(status.alloc((size_t)ntiles * RS_BINS));
    CSB_TRY(ticket.alloc(RS_MAX_PASSES));
    CSB_TRY(keys_sorted.alloc((size_t)nnz));
    if (npasses >= 2) CSB_TRY(bufA.alloc((size_t)nnz * recsz));
    if (npasses >= 3) CSB_TRY(bufB.alloc((size_t)nnz * recsz));
    RS_CUDA(cudaMemsetAsync(hist.ptr, 0, RS_MAX_PASSES * RS_BINS * sizeof(unsigned long long), s));
    RS_CUDA(cudaMemsetAsync(ticket.ptr, 0, RS_MAX_PASSES * sizeof(unsigned), s));
    k_rs_hist<<<min(ceil_div(nnz, 1024 * 8), sm_count() * 8), 256, 0, s>>>(key, nnz, npasses, hist.ptr);
    CSB_LAUNCHED();
    k_rs_starts<<<1, RS_BINS, 0, s>>>(npasses, hist.ptr);
    CSB_LAUNCHED();
    const int src0 = a ? 0 : 2;
    for (int q = 0; q < npasses; q++) {
        RS_CUDA(cudaMemsetAsync(status.ptr, 0, (size_t)ntiles * RS_BINS * sizeof(unsigned long long), s));
        const bool first = q == 0, last = q == npasses - 1;
        const void *rin = first ? nullptr : ((q & 1) ? bufA.ptr : bufB.ptr);
        void *rout = last ? nullptr : ((q & 1) ? bufB.ptr : bufA.ptr);
        const unsigned long long *ds = hist.ptr + q * RS_BINS;
#define RS_LAUNCH(SRC, DST)                                                                       \
        do {                                                                                      \
            if (staged) {                                                                         \
                RS_CUDA(cudaFuncSetAttribute(k_rs_pass_s<SRC, DST, VALUES>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs2_smem<VALUES>())); \
                k_rs_pass_s<SRC, DST, VALUES><<<ntiles, RS2_THREADS, rs2_smem<VALUES>(), s>>>(nnz, 8 * q, key, a, v, rin, Ap, ncols, \
                    keys_sorted.ptr, a_out, v_out, rout, ds, status.ptr, ticket.ptr + q, lb);         \
            } else {                                                                              \
                RS_CUDA(cudaFuncSetAttribute(k_rs_pass<SRC, DST, VALUES>, cudaFuncAttributeMaxDynamicSharedMemorySize, rs_pass_smem(SRC))); \
                k_rs_pass<SRC, DST, VALUES><<<ntiles, RS_THREADS, rs_pass_smem(SRC), s>>>(nnz, 8 * q, key, a, v, rin, Ap, ncols, \
                    keys_sorted.ptr, a_out, v_out, rout, ds, status.ptr, ticket.ptr + q, lb);         \
            }                                                                                     \
        } while (0)
        if (first && last)      { if (src0 == 0) RS_LAUNCH(0, 0); else RS_LAUNCH(2, 0); }
        else if (first)         { if (src0 == 0) RS_LAUNCH(0, 1); else RS_LAUNCH(2, 1); }
        else if (last)          RS_LAUNCH(1, 0);
        else                    RS_LAUNCH(1, 1);
#undef RS_LAUNCH
        CSB_LAUNCHED();
    }
    k_rs_bounds<<<ceil_div((long long)nkeys + 1, 256), 256, 0, s>>>(keys_sorted.ptr, nnz, nkeys, Cp);
    CSB_LAUNCHED();
    return CSB200_OK;
}

// Sorts nnz > 0 triples (key[e], a[e], v[e]) stably by key; a == nullptr derives the int
// payload from Ap (ncols columns) as the column holding position e; v == nullptr sorts
// pattern only.  Outputs a_out / v_out (nnz) and Cp (nkeys + 1).  All device pointers.
int stable_sort_by_key(long long nnz, int nkeys, const int *key, const int *a, const csi *Ap, int ncols,
                       const double *v, csi *Cp, csi *a_out, double *v_out)
{
    if (v) return sort_impl<true>(nnz, nkeys, key, a, Ap, ncols, v, Cp, a_out, v_out);
    return sort_impl<false>(nnz, nkeys, key, a, Ap, ncols, nullptr, Cp, a_out, nullptr);
}

}  // namespace csb
