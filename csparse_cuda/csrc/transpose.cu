// transpose.cu -- cs_transpose (csparse.py:2292-2315): C = A' as a counting sort by
// row index.  The reference is a sequential stable scatter (histogram :2305-2306,
// cs_cumsum :2307, scatter :2308-2314); the result must be bit-identical (p, i, x),
// so the order inside every output column has to be the source order.
//
// Per-entry global atomics and 4/8-byte scattered stores cost one L2 transaction
// each and capped a direct scatter at 13 % of the HBM roofline on B200 (measured,
// profiles/r1_probe_launches.csv).  This version moves the data in two coalesced
// hops instead -- a two-level (bucket, row) counting sort:
//
//   k_bucket_hist   rows are grouped into buckets of RB consecutive rows; count
//                   entries per bucket (warp-aggregated: ~4 atomics per 32 entries)
//   excl_scan       bucket offsets (scan.cu) -- also the bucket fill cursors
//   k_tile_cols     column of the first entry of every 4096-entry tile
//   k_partition     stream (Ai, Ax), attach the column id, append one packed word
//                   (row inside the bucket << colbits | source column) and the value
//                   to the entry's bucket (one atomic per (tile, bucket) reserves the
//                   slots); matrices whose packed word would need more than 31 bits
//                   take the radix path
//   k_bucket_sort   one CTA per bucket: count rows in shared memory (this IS the
//                   reference's histogram + cs_cumsum, restricted to RB rows),
//                   scatter the bucket into a shared-memory staging area, put every
//                   row into source order, write Cp / Ci / Cx fully coalesced
//   k_bucket_big    buckets larger than the staging area (power-law rows): the same
//                   steps with the staging area in global memory
//
// Source order: entries of one output column come from distinct source columns (or
// are duplicates of one (i,j) pair), so source order == ascending j with ties in
// storage order.  Buckets are filled in whatever order the atomics land; each row
// is then sorted by j, and tied groups are re-read from the source column in
// storage order.  Correctness never depends on how atomics were ordered.
#include "common.cuh"
#include <algorithm>
#include <stdlib.h>
#include <type_traits>
#include <vector>

namespace csb {

constexpr int TR_THREADS = 256;
constexpr int TR_TILE = 4096;                 // entries per histogram CTA
constexpr int PT_EPT = 8;                     // entries per thread in the partition kernel
constexpr int PT_TILE = TR_THREADS * PT_EPT;  // 2048 entries per partition CTA
constexpr int PT_SMEM_COLS = 3072;            // column pointers staged per partition tile
constexpr int HIST_WIN = 4096;                // bucket window counted in shared memory
constexpr int WB_EPT = 12;                    // entries per lane in the warp-per-bucket kernel
constexpr int WB_CAP = 32 * WB_EPT;           // 384 entries staged per warp
constexpr int WB_RB_MAX = 128;                // rows per bucket (power of two)
constexpr int WB_WARPS = 8;                   // warps (independent bucket pipelines) per CTA
constexpr int WB_WARP_BYTES = WB_CAP * 4 + (2 * WB_RB_MAX + 8) * 4 + WB_CAP * 2;   // scol, cnt, start, sidx
constexpr int BK_THREADS = 256;
constexpr int BK_EPT = 12;                    // entries per thread
constexpr int BK_CAP = BK_THREADS * BK_EPT;   // 3072 entries staged per bucket
constexpr int BK_RB_MAX = 1024;               // rows per bucket (power of two)
constexpr int BK_SMEM = BK_CAP * 12 + 2 * (BK_RB_MAX + 1) * 4 + 64;
constexpr int FIX_SHORT = 32;                 // rows up to this length: one thread (global-memory path)
#ifndef FIX_THREAD_DEF
#define FIX_THREAD_DEF 8
#endif
constexpr int FIX_THREAD = FIX_THREAD_DEF;                 // staged rows up to this length: one thread; up to 32: warp rank sort


// ---- tile -> first column ------------------------------------------------------
__global__ void k_tile_cols(const csi *__restrict__ Ap, int n, long long nnz, int ntiles,
                            int *__restrict__ tile_col)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == ntiles) { tile_col[t] = n - 1; return; }
    const long long p0 = (long long)t * PT_TILE;
    tile_col[t] = upper_row(Ap, 0, n, (int)p0);   // largest j with Ap[j] <= p0
}

// block-wide min / max of a per-thread value (256 threads)
__device__ __forceinline__ void block_minmax(int lo, int hi, int *s_red, int &bmin, int &bmax)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { s_red[wid] = lo; s_red[8 + wid] = hi; }
    __syncthreads();
    bmin = s_red[0]; bmax = s_red[8];
#pragma unroll
    for (int w = 1; w < TR_THREADS / 32; w++) { bmin = min(bmin, s_red[w]); bmax = max(bmax, s_red[8 + w]); }
    __syncthreads();
}

// ---- bucket histogram ------------------------------------------------------------
// A tile of consecutive entries touches a narrow window of buckets for banded
// matrices: count in a shared-memory window and flush one global atomic per
// (tile, bucket).  Tiles whose window does not fit are only counted (wide_tiles):
// such a matrix is transposed by the stable radix sort of radix.cu.
__global__ void __launch_bounds__(TR_THREADS)
k_bucket_hist(const csi *__restrict__ Ai, long long nnz, int log_rb, int *__restrict__ bcount,
              int *__restrict__ wide_tiles, int *__restrict__ tile_range)
{
    __shared__ int cnt[HIST_WIN];
    __shared__ int s_red[16];
    const long long p_begin = (long long)blockIdx.x * TR_TILE;
    const long long p_end = min(nnz, p_begin + TR_TILE);
    int b[16];
    int lo = INT_MAX, hi = -1;
#pragma unroll
    for (int k = 0; k < 4; k++) {                       // four independent 16-byte loads in flight
        const long long p = p_begin + (long long)(k * TR_THREADS + threadIdx.x) * 4;
        if (p + 3 < p_end) {
            const int4 r = ldg_stream(reinterpret_cast<const int4 *>(Ai + p));
            b[4 * k] = r.x >> log_rb; b[4 * k + 1] = r.y >> log_rb; b[4 * k + 2] = r.z >> log_rb; b[4 * k + 3] = r.w >> log_rb;
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) b[4 * k + e] = p + e < p_end ? (Ai[p + e] >> log_rb) : -1;
        }
    }
#pragma unroll
    for (int k = 0; k < 16; k++) if (b[k] >= 0) { lo = min(lo, b[k]); hi = max(hi, b[k]); }
    int bmin, bmax;
    block_minmax(lo, hi, s_red, bmin, bmax);
    if (bmax < 0) return;
    // first and last bucket this tile feeds: the host schedules the partition / sort slabs from it
    if (threadIdx.x == 0) { tile_range[2 * blockIdx.x] = bmin; tile_range[2 * blockIdx.x + 1] = bmax; }
    const int win = bmax - bmin + 1;
    if (win <= HIST_WIN) {
        for (int k = threadIdx.x; k < win; k += TR_THREADS) cnt[k] = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; k++) if (b[k] >= 0) atomicAdd(&cnt[b[k] - bmin], 1);
        __syncthreads();
        for (int k = threadIdx.x; k < win; k += TR_THREADS)
            if (cnt[k]) atomicAdd(&bcount[bmin + k], cnt[k]);
    } else if (threadIdx.x == 0) {
        // rows all over the matrix: the bucket sort would need one global atomic per entry here
        // and in the partition; the caller switches to the radix sort instead
        atomicAdd(wide_tiles, 1);
    }
}

// ---- partition into buckets ---------------------------------------------------------
// Tile of PT_TILE consecutive entries, all loaded up front (8 per thread).  Ranks
// inside the tile come from shared-memory counters over the tile's bucket window;
// one global atomic per (tile, bucket) reserves the slots, so the global atomics
// are few and all in flight together.
// `arrived` (fused kernel only): per bucket, how many of its entries have been WRITTEN; the tile that
// brings a bucket to its full size lists it in ready[] (shared memory) -- that bucket can be sorted now,
// and its share of the intermediate is still in L2.
template <bool VALUES>
__device__ __forceinline__ void
partition_tile(const int t, const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
               long long nnz, const int *__restrict__ tile_col, int log_rb, int colbits, int *__restrict__ bfill,
               int *__restrict__ ikey, double *__restrict__ ival,
               int *__restrict__ arrived, const int *__restrict__ bstart, int *ready, int *nready)
{
    __shared__ int sAp[PT_SMEM_COLS];
    __shared__ int cnt[HIST_WIN];
    __shared__ unsigned short cnt_own[HIST_WIN];   // fused kernel: this tile's entries per bucket (<= PT_TILE; cnt becomes the base)
    __shared__ int s_red[16];
    const long long p_begin = (long long)t * PT_TILE;
    const long long p_end = min(nnz, p_begin + PT_TILE);
    const int j_first = tile_col[t];
    const int j_last = tile_col[t + 1];           // >= column of the last entry of this tile
    const int ncols = j_last - j_first + 1;
    const bool staged = ncols + 1 <= PT_SMEM_COLS;
    if (staged)
        for (int k = threadIdx.x; k <= ncols; k += TR_THREADS) sAp[k] = Ap[j_first + k];

    int rows[PT_EPT], cols[PT_EPT], rank[PT_EPT];
    double vals[PT_EPT];
    int cntk[PT_EPT / 4];
#pragma unroll
    for (int k = 0; k < PT_EPT / 4; k++) {
        const long long p = p_begin + (long long)(k * TR_THREADS + threadIdx.x) * 4;
        cntk[k] = p < p_end ? (int)min((long long)4, p_end - p) : 0;
        if (cntk[k] == 4) {
            const int4 r = ldg_stream(reinterpret_cast<const int4 *>(Ai + p));
            rows[4 * k] = r.x; rows[4 * k + 1] = r.y; rows[4 * k + 2] = r.z; rows[4 * k + 3] = r.w;
            if (VALUES) {
                ldg_stream4(Ax + p, vals[4 * k], vals[4 * k + 1], vals[4 * k + 2], vals[4 * k + 3]);
            }
        } else {
#pragma unroll
            for (int e = 0; e < 4; e++) {
                rows[4 * k + e] = e < cntk[k] ? Ai[p + e] : -1;
                if (VALUES) vals[4 * k + e] = e < cntk[k] ? Ax[p + e] : 0.0;
            }
        }
    }
    int lo = INT_MAX, hi = -1;
#pragma unroll
    for (int k = 0; k < PT_EPT; k++)
        if (k % 4 < cntk[k / 4]) { const int b = rows[k] >> log_rb; lo = min(lo, b); hi = max(hi, b); }
    int bmin, bmax;
    block_minmax(lo, hi, s_red, bmin, bmax);      // also orders the sAp staging before its use
    const int win = bmax - bmin + 1;
    const bool windowed = win <= HIST_WIN;
    if (windowed) {
        for (int k = threadIdx.x; k < win; k += TR_THREADS) cnt[k] = 0;
        __syncthreads();
    }
    // column of every entry + rank inside the tile's share of its bucket
#pragma unroll
    for (int k = 0; k < PT_EPT / 4; k++) {
        if (cntk[k] > 0) {
            const int p = (int)(p_begin + (long long)(k * TR_THREADS + threadIdx.x) * 4);
            int j;   // largest j with Ap[j] <= p
            if (staged) j = upper_row(sAp, 0, ncols - 1, p);
            else        j = upper_row(Ap, j_first, j_last, p) - j_first;
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (e < cntk[k]) {
                    const int pe = p + e;
                    if (staged) { while (sAp[j + 1] <= pe) j++; }
                    else        { while (Ap[j_first + j + 1] <= pe) j++; }
                    cols[4 * k + e] = j_first + j;
                    const int b = rows[4 * k + e] >> log_rb;
                    rank[4 * k + e] = windowed ? atomicAdd(&cnt[b - bmin], 1) : atomicAdd(&bfill[b], 1);
                }
            }
        }
    }
    if (windowed) {
        __syncthreads();
        for (int k = threadIdx.x; k < win; k += TR_THREADS) {
            const int c = cnt[k];
            if (arrived) cnt_own[k] = (unsigned short)c;
            if (c) cnt[k] = atomicAdd(&bfill[bmin + k], c);      // cnt becomes the reserved base
        }
        __syncthreads();
    }
#pragma unroll
    for (int k = 0; k < PT_EPT; k++) {
        if (k % 4 < cntk[k / 4]) {
            const int b = rows[k] >> log_rb;
            const long long pos = (long long)rank[k] + (windowed ? cnt[b - bmin] : 0);
            // one word per entry: (row inside the bucket, source column)
            ikey[pos] = ((rows[k] & ((1 << log_rb) - 1)) << colbits) | cols[k];
            if (VALUES) ival[pos] = vals[k];
        }
    }
    if (arrived) {                                   // windowed is guaranteed by the caller (no wide tiles)
        __threadfence();                             // this tile's entries are visible before they are counted
        __syncthreads();
        for (int k = threadIdx.x; k < win; k += TR_THREADS) {
            const int c = cnt_own[k];
            if (c) {
                const int b = bmin + k;
                const int before = atomicAdd(&arrived[b], c);
                if (before + c == bstart[b + 1] - bstart[b]) ready[atomicAdd(nready, 1)] = b;
            }
        }
        __syncthreads();
    }
}

template <bool VALUES>
__global__ void __launch_bounds__(TR_THREADS, 4)
k_partition(const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
            long long nnz, const int *__restrict__ tile_col, int log_rb, int colbits, int *__restrict__ bfill,
            int *__restrict__ ikey, double *__restrict__ ival, int tile0)
{
    partition_tile<VALUES>(tile0 + (int)blockIdx.x, Ap, Ai, Ax, nnz, tile_col, log_rb, colbits, bfill, ikey, ival,
                           nullptr, nullptr, nullptr, nullptr);
}

// ---- order repair helpers ------------------------------------------------------------
// After sorting an output column by j, runs of equal j are duplicates of one
// (row r, column j) entry of A; their values must appear in A's storage order.
__device__ void fix_tied_group(const csi *Ap, const csi *Ai, const double *Ax,
                               int r, int j, double *cx, int g)
{
    int k = 0;
    for (int p = Ap[j]; p < Ap[j + 1] && k < g; p++)
        if (Ai[p] == r) cx[k++] = Ax[p];
}

// one thread: in-place insertion sort of a short row (len <= FIX_SHORT) by column
template <bool VALUES>
__device__ void thread_fix_row(int r, csi *ci, double *cx, int len,
                               const csi *Ap, const csi *Ai, const double *Ax)
{
    bool sorted = true;
    for (int k = 1; k < len; k++) sorted &= ci[k] > ci[k - 1];
    if (sorted) return;
    for (int a = 1; a < len; a++) {
        const int kj = ci[a];
        const double kv = VALUES ? cx[a] : 0.0;
        int c = a - 1;
        while (c >= 0 && ci[c] > kj) {
            ci[c + 1] = ci[c];
            if (VALUES) cx[c + 1] = cx[c];
            c--;
        }
        ci[c + 1] = kj;
        if (VALUES) cx[c + 1] = kv;
    }
    if (VALUES) {
        for (int k = 0; k + 1 < len;) {
            int g = 1;
            while (k + g < len && ci[k + g] == ci[k]) g++;
            if (g > 1) fix_tied_group(Ap, Ai, Ax, r, ci[k], cx + k, g);
            k += g;
        }
    }
}

// Cooperative in-place sort of one row by a group of G threads (G = 32: a warp,
// otherwise the whole CTA).  Normalised bitonic network: every comparator moves the
// smaller key to the lower index, so virtual +inf padding above `len` never moves
// and any length works.  Works on shared or global memory.
template <int G, bool VALUES, class V = double>
__device__ void group_sort_row(csi *ci, V *cx, int len, int tid)
{
    int P = 2;
    while (P < len) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += G) {
                const int lo = ((t / j) * (j << 1)) + (t % j);
                const int hi = (j == (k >> 1)) ? (lo ^ (k - 1)) : (lo + j);
                if (hi < len) {
                    const int a = ci[lo], b = ci[hi];
                    if (a > b) {
                        ci[lo] = b; ci[hi] = a;
                        if (VALUES) { const V xa = cx[lo]; cx[lo] = cx[hi]; cx[hi] = xa; }
                    }
                }
            }
            if (G == 32) __syncwarp(); else __syncthreads();
        }
    }
}

template <int G, bool VALUES>
__device__ bool group_fix_row(int r, csi *ci, double *cx, int len,
                              const csi *Ap, const csi *Ai, const double *Ax, int tid, int *flag)
{
    bool bad = false;
    for (int t = tid; t + 1 < len; t += G) bad |= ci[t] >= ci[t + 1];
    if (G == 32) {
        bad = __any_sync(0xffffffffu, bad);
    } else {
        if (tid == 0) *flag = 0;
        __syncthreads();
        if (bad) *flag = 1;
        __syncthreads();
        bad = *flag != 0;
        __syncthreads();
    }
    if (!bad) return false;
    group_sort_row<G, VALUES>(ci, cx, len, tid);
    if (VALUES) {
        for (int t = tid; t + 1 < len; t += G) {
            if (ci[t] == ci[t + 1] && (t == 0 || ci[t - 1] != ci[t])) {
                int g = 2;
                while (t + g < len && ci[t + g] == ci[t]) g++;
                fix_tied_group(Ap, Ai, Ax, r, ci[t], cx + t, g);
            }
        }
        if (G == 32) __syncwarp(); else __syncthreads();
    }
    return true;
}

// One warp sorts a row of 9..32 entries by counting: every lane ranks its own entry
// against the others (shuffles only), then all entries move at once.  Stable.
template <bool HAS_V, class V>
__device__ __forceinline__ bool warp_rank_sort(csi *ci, V *cv, int len, int lane)
{
    const int c = lane < len ? ci[lane] : INT_MAX;
    V v = V();
    if (HAS_V && lane < len) v = cv[lane];
    // Keys below 2^26 (any matrix of < 67 M columns): a bitonic network on (key << 5 | lane), 15
    // compare-exchange steps of one shuffle each instead of `len` rounds of shuffle + two compares
    // (the rank loop was 45 % of k_bucket_sort's stall samples on the 27-point stencil); the low
    // bits keep equal keys in their order of arrival and name the lane whose value follows.
    if (__all_sync(0xffffffffu, lane >= len || (unsigned)c < (1u << 26))) {
        unsigned w = lane < len ? ((unsigned)c << 5) | (unsigned)lane : 0xffffffffu;
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const unsigned o = __shfl_xor_sync(0xffffffffu, w, j);
                const bool keep_min = ((lane & j) == 0) == ((lane & k) == 0);
                w = keep_min ? min(w, o) : max(w, o);
            }
        }
        const unsigned wprev = __shfl_up_sync(0xffffffffu, w, 1);
        const bool tie2 = lane > 0 && lane < len && (wprev >> 5) == (w >> 5);
        V vv = V();
        if (HAS_V) {
            if constexpr (sizeof(V) == 8) vv = __shfl_sync(0xffffffffu, v, (int)(w & 31u));
            else vv = (V)__shfl_sync(0xffffffffu, (int)v, (int)(w & 31u));
        }
        __syncwarp();
        if (lane < len) { ci[lane] = (int)(w >> 5); if (HAS_V) cv[lane] = vv; }
        __syncwarp();
        return __any_sync(0xffffffffu, tie2);
    }
    int rank = 0;
    bool tie = false;
    for (int u = 0; u < len; u++) {
        const int cu = __shfl_sync(0xffffffffu, c, u);
        rank += (cu < c) || (cu == c && u < lane);
        tie |= cu == c && u != lane;
    }
    __syncwarp();
    if (lane < len) { ci[rank] = c; if (HAS_V) cv[rank] = v; }
    __syncwarp();
    return __any_sync(0xffffffffu, tie && lane < len);
}

// values of tied (duplicate) groups of a sorted row, re-read from A in storage order
__device__ __forceinline__ void warp_fix_ties(int r, const csi *ci, double *cx, int len,
                                              const csi *Ap, const csi *Ai, const double *Ax, int lane)
{
    for (int t = lane; t + 1 < len; t += 32) {
        if (ci[t] == ci[t + 1] && (t == 0 || ci[t - 1] != ci[t])) {
            int g = 2;
            while (t + g < len && ci[t + g] == ci[t]) g++;
            fix_tied_group(Ap, Ai, Ax, r, ci[t], cx + t, g);
        }
    }
    __syncwarp();
}

// exclusive scan of cnt[0..n) (n <= BK_RB_MAX) into start[0..n], by the whole CTA
__device__ void block_scan_rows(const int *cnt, int *start, int n, int *warp_tot)
{
    constexpr int PER = BK_RB_MAX / BK_THREADS;      // 4 consecutive rows per thread
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int v[PER], s = 0;
#pragma unroll
    for (int k = 0; k < PER; k++) { const int idx = tid * PER + k; v[k] = idx < n ? cnt[idx] : 0; s += v[k]; }
    int inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int w = lane < BK_THREADS / 32 ? warp_tot[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, winc, o); if (lane >= o) winc += t; }
        if (lane < BK_THREADS / 32) warp_tot[lane] = winc - w;
    }
    __syncthreads();
    int e = warp_tot[wid] + inc - s;
#pragma unroll
    for (int k = 0; k < PER; k++) { const int idx = tid * PER + k; if (idx <= n) start[idx] = e; e += v[k]; }
    __syncthreads();
}

// ---- one CTA per bucket, staging in shared memory ---------------------------------------
template <bool VALUES>
__global__ void __launch_bounds__(BK_THREADS, 5)
k_bucket_sort(int m, int log_rb, int colbits, int nbuckets, const int *__restrict__ bstart,
              const int *__restrict__ ikey, const double *__restrict__ ival,
              const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
              csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx, int b0)
{
    extern __shared__ __align__(16) unsigned char smem[];
    double *sval = reinterpret_cast<double *>(smem);
    int *scol = reinterpret_cast<int *>(smem + BK_CAP * 8);
    int *rowcnt = scol + BK_CAP;
    int *rowstart = rowcnt + BK_RB_MAX + 1;
    int *warp_tot = rowstart + BK_RB_MAX + 1;          // 16 ints
    const int b = b0 + blockIdx.x;
    const int tid = threadIdx.x;
    const int rb = 1 << log_rb;
    const int R0 = b << log_rb;
    const int nrows = min(rb, m - R0);
    const int base = bstart[b];
    const int nb = bstart[b + 1] - base;
    if (nb > BK_CAP) return;                           // k_bucket_big's job

    for (int k = tid; k <= rb; k += BK_THREADS) rowcnt[k] = 0;
    __syncthreads();
    const int colmask = (1 << colbits) - 1;
    int rank[BK_EPT];
#pragma unroll
    for (int k = 0; k < BK_EPT; k++) {
        const int e = tid + k * BK_THREADS;
        rank[k] = 0;
        if (e < nb) rank[k] = atomicAdd(&rowcnt[ikey[base + e] >> colbits], 1);
    }
    __syncthreads();
    block_scan_rows(rowcnt, rowstart, nrows, warp_tot);
#pragma unroll
    for (int k = 0; k < BK_EPT; k++) {
        const int e = tid + k * BK_THREADS;
        if (e < nb) {
            const int key = ikey[base + e];                  // coalesced re-read (L1 / L2 hit)
            const int pos = rowstart[key >> colbits] + rank[k];
            scol[pos] = key & colmask;
            if (VALUES) sval[pos] = ival[base + e];
        }
    }
    __syncthreads();
    // source order inside every row: <= 8 entries by one thread, <= 32 by a warp's rank sort,
    // longer rows by a warp's bitonic network
    bool any_long = false;
    for (int rl = tid; rl < nrows; rl += BK_THREADS) {
        const int s = rowstart[rl], len = rowstart[rl + 1] - s;
        if (len > FIX_THREAD) any_long = true;
        else if (len > 1) thread_fix_row<VALUES>(R0 + rl, scol + s, sval + s, len, Ap, Ai, Ax);
    }
    if (__syncthreads_or(any_long)) {
        const int lane = tid & 31;
        for (int rl = tid >> 5; rl < nrows; rl += BK_THREADS / 32) {
            const int s = rowstart[rl], len = rowstart[rl + 1] - s;
            if (len > FIX_SHORT) {
                group_fix_row<32, VALUES>(R0 + rl, scol + s, sval + s, len, Ap, Ai, Ax, lane, nullptr);
            } else if (len > FIX_THREAD) {
                if (warp_rank_sort<VALUES, double>(scol + s, sval + s, len, lane) && VALUES)
                    warp_fix_ties(R0 + rl, scol + s, sval + s, len, Ap, Ai, Ax, lane);
            }
        }
        __syncthreads();
    }
    // coalesced output: this bucket owns C's columns R0..R0+nrows and the slots [base, base+nb)
    for (int rl = tid; rl < nrows; rl += BK_THREADS) Cp[R0 + rl] = base + rowstart[rl];
    if (b == nbuckets - 1 && tid == 0) Cp[m] = base + nb;
    for (int t = tid; t < nb; t += BK_THREADS) {
        Ci[base + t] = scol[t];
        if (VALUES) Cx[base + t] = sval[t];
    }
}

// ---- one WARP per bucket, staging in the warp's slice of shared memory ---------------------
// Buckets are small (<= WB_CAP entries, <= WB_RB_MAX rows) so that a single warp can count,
// scan, scatter and order one bucket with __syncwarp only: the warps of an SM run as many
// independent pipelines.  Each warp walks its own sequence of buckets and has the row fields
// of the NEXT bucket in flight while it works on the current one.  Only (column, source slot)
// pairs are staged; values are gathered from the bucket (L1/L2-resident) on output.
// One warp, a sequence of buckets bucket_at(first), bucket_at(first + stride), ... < count.  CG: the
// intermediate was written by other CTAs of this same launch (fused kernel): read it from L2.
template <bool VALUES, bool CG, class BucketAt>
__device__ __forceinline__ void
warp_sort_buckets(const int first, const int stride, const int count, BucketAt bucket_at, unsigned char *mine,
                  int m, int log_rb, int colbits, int nbuckets, const int *__restrict__ bstart,
                  const int *ikey, const double *ival,
                  const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
                  csi *__restrict__ Cp, csi *Ci, double *Cx)
{
    auto LDK = [](const int *q) { return CG ? __ldcg(q) : *q; };
    auto LDV = [](const double *q) { return CG ? __ldcg(q) : *q; };
    const int lane = threadIdx.x & 31;
    int *scol = reinterpret_cast<int *>(mine);
    int *cnt = scol + WB_CAP;                      // WB_RB_MAX
    int *start = cnt + WB_RB_MAX;                  // WB_RB_MAX + 1
    unsigned short *sidx = reinterpret_cast<unsigned short *>(start + WB_RB_MAX + 8);
    const int rb = 1 << log_rb;
    const int colmask = (1 << colbits) - 1;

    int idx = first;                                   // position in the sequence of buckets bucket_at(0 .. count)
    if (idx >= count) return;
    int b = bucket_at(idx);
    // prologue: bounds and row fields of the first bucket
    int base = bstart[b], nb = bstart[b + 1] - base;
    int rows[WB_EPT];
#pragma unroll
    for (int k = 0; k < WB_EPT; k++) {
        const int e = lane + k * 32;
        rows[k] = (e < nb && nb <= WB_CAP) ? LDK(ikey + base + e) : 0;      // packed (local row, column) words
    }
    while (idx < count) {
        const int R0 = b << log_rb;
        const int nrows = min(rb, m - R0);
        const int idxn = idx + stride;             // next bucket of this warp
        const int bn = idxn < count ? bucket_at(idxn) : 0;
        int nbase = 0, nnb = 0;
        if (idxn < count) { nbase = bstart[bn]; nnb = bstart[bn + 1] - nbase; }

        if (nb <= WB_CAP) {                        // larger buckets are k_bucket_big's job
            for (int k = lane; k < WB_RB_MAX; k += 32) cnt[k] = 0;
            __syncwarp();
            int slot[WB_EPT];                      // (local row << 16) | rank inside the row
#pragma unroll
            for (int k = 0; k < WB_EPT; k++) {
                const int rl = rows[k] >> colbits;
                slot[k] = (rl << 16) | ((lane + k * 32 < nb) ? atomicAdd(&cnt[rl], 1) : 0);
            }
            // the next bucket's row fields go in flight now
#pragma unroll
            for (int k = 0; k < WB_EPT; k++) {
                const int e = lane + k * 32;
                rows[k] = (e < nnb && nnb <= WB_CAP) ? LDK(ikey + nbase + e) : 0;
            }
            __syncwarp();
            {   // exclusive scan of cnt[0..WB_RB_MAX) -> start[0..WB_RB_MAX]; 4 consecutive rows per lane
                constexpr int PER = WB_RB_MAX / 32;
                int v[PER], sum = 0;
#pragma unroll
                for (int k = 0; k < PER; k++) { v[k] = cnt[lane * PER + k]; sum += v[k]; }
                int inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
                int e = inc - sum;
#pragma unroll
                for (int k = 0; k < PER; k++) { start[lane * PER + k] = e; e += v[k]; }
                if (lane == 31) start[WB_RB_MAX] = e;
            }
            __syncwarp();
#pragma unroll
            for (int k = 0; k < WB_EPT; k++) {
                const int e = lane + k * 32;
                if (e < nb) {
                    const int pos = start[slot[k] >> 16] + (slot[k] & 0xffff);
                    scol[pos] = LDK(ikey + base + e) & colmask;            // coalesced re-read (L1 hit)
                    sidx[pos] = (unsigned short)e;
                }
            }
            __syncwarp();
            // source order inside every row: insertion sort of (column, slot) pairs
            bool any_long = false, any_tie = false;
            for (int r = lane; r < nrows; r += 32) {
                const int s0 = start[r], len = start[r + 1] - s0;
                if (len > FIX_THREAD) { any_long = true; continue; }
                int *ci = scol + s0;
                unsigned short *ix = sidx + s0;
                for (int a = 1; a < len; a++) {
                    const int kj = ci[a];
                    const unsigned short kv = ix[a];
                    int c = a - 1;
                    while (c >= 0 && ci[c] > kj) { ci[c + 1] = ci[c]; ix[c + 1] = ix[c]; c--; }
                    ci[c + 1] = kj; ix[c + 1] = kv;
                    any_tie |= c >= 0 && ci[c] == kj;
                }
            }
            __syncwarp();
            if (__any_sync(0xffffffffu, any_long)) {
                for (int r = 0; r < nrows; r++) {
                    const int s0 = start[r], len = start[r + 1] - s0;
                    if (len > FIX_SHORT) {
                        group_sort_row<32, true, unsigned short>(scol + s0, sidx + s0, len, lane);
                        for (int t = lane; t + 1 < len; t += 32) any_tie |= scol[s0 + t] == scol[s0 + t + 1];
                    } else if (len > FIX_THREAD) {
                        any_tie |= warp_rank_sort<true, unsigned short>(scol + s0, sidx + s0, len, lane);
                    }
                }
                __syncwarp();
            }
            for (int r = lane; r < nrows; r += 32) Cp[R0 + r] = base + start[r];
            if (b == nbuckets - 1 && lane == 0) Cp[m] = base + nb;
            for (int t = lane; t < nb; t += 32) {
                Ci[base + t] = scol[t];
                if (VALUES) Cx[base + t] = LDV(ival + base + sidx[t]);
            }
            if (VALUES && __any_sync(0xffffffffu, any_tie)) {
                // duplicates of one (i,j) pair: their values go out in A's storage order
                __syncwarp();
                for (int r = lane; r < nrows; r += 32) {
                    const int s0 = start[r], len = start[r + 1] - s0;
                    for (int k = 0; k + 1 < len;) {
                        int g = 1;
                        while (k + g < len && scol[s0 + k + g] == scol[s0 + k]) g++;
                        if (g > 1) fix_tied_group(Ap, Ai, Ax, R0 + r, scol[s0 + k], Cx + base + s0 + k, g);
                        k += g;
                    }
                }
            }
            __syncwarp();
        } else {
#pragma unroll
            for (int k = 0; k < WB_EPT; k++) {
                const int e = lane + k * 32;
                rows[k] = (e < nnb && nnb <= WB_CAP) ? LDK(ikey + nbase + e) : 0;
            }
        }
        idx = idxn; b = bn; base = nbase; nb = nnb;
    }
}


template <bool VALUES>
__global__ void __launch_bounds__(WB_WARPS * 32, 4)
k_bucket_sort_warp(int m, int log_rb, int colbits, int nbuckets, const int *__restrict__ bstart,
                   const int *__restrict__ ikey, const double *__restrict__ ival,
                   const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
                   csi *__restrict__ Cp, csi *Ci, double *Cx, int b0, int b1)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const int wid = threadIdx.x >> 5;
    // this launch sorts the buckets [b0, b1), strided over all warps of the grid
    warp_sort_buckets<VALUES, false>(blockIdx.x * WB_WARPS + wid, gridDim.x * WB_WARPS, b1 - b0,
                                     [b0](int i) { return b0 + i; }, smem + wid * WB_WARP_BYTES,
                                     m, log_rb, colbits, nbuckets, bstart, ikey, ival, Ap, Ai, Ax, Cp, Ci, Cx);
}

// ---- partition and sort in ONE persistent launch (warp-per-bucket matrices) -----------------------
// Tiles are claimed in order through a ticket.  After a CTA has written a tile's entries it adds them
// to the buckets' arrival counters; the tile that completes a bucket lists it, and the CTA's eight warps
// sort the listed buckets at once -- their share of the intermediate was written moments ago and is read
// back from L2, not HBM.  Nothing waits for anything: a bucket is sorted by whoever completes it.
template <bool VALUES>
__global__ void __launch_bounds__(TR_THREADS, 3)
k_partition_sort(const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
                 long long nnz, int ntiles, const int *__restrict__ tile_col, int m, int log_rb, int colbits, int nbuckets,
                 int *__restrict__ bfill, const int *__restrict__ bstart, int *__restrict__ arrived, unsigned *ticket,
                 int *ikey, double *ival, csi *__restrict__ Cp, csi *Ci, double *Cx)
{
    extern __shared__ __align__(16) unsigned char smem[];        // WB_WARPS x WB_WARP_BYTES
    __shared__ int ready[PT_TILE];                               // a tile completes at most one bucket per entry
    __shared__ int nready;
    __shared__ unsigned s_tile;
    const int wid = threadIdx.x >> 5;
    while (true) {
        if (threadIdx.x == 0) { s_tile = atomicAdd(ticket, 1u); nready = 0; }
        __syncthreads();
        const unsigned t = s_tile;
        if (t >= (unsigned)ntiles) break;
        partition_tile<VALUES>((int)t, Ap, Ai, Ax, nnz, tile_col, log_rb, colbits, bfill, ikey, ival,
                               arrived, bstart, ready, &nready);
        __threadfence();                                        // acquire side of the arrival counters
        const int nr = nready;
        if (nr > 0)
            warp_sort_buckets<VALUES, true>(wid, WB_WARPS, nr, [&](int i) { return ready[i]; }, smem + wid * WB_WARP_BYTES,
                                            m, log_rb, colbits, nbuckets, bstart, ikey, ival, Ap, Ai, Ax, Cp, Ci, Cx);
        __syncthreads();                                        // ready[] is reused by the next tile
    }
}

// ---- buckets that do not fit shared memory (power-law rows) ---------------------------------
// One CTA sorts the whole bucket by the composite key (local row, source column) with a
// stable LSD radix sort in global memory: an odd number of passes ping-pongs between the
// bucket's slice of the partition buffer and a scratch slice, the last pass writes Ci / Cx.
// Cost is linear in the bucket size whatever the row lengths are (one row may hold the
// whole bucket).  Each warp owns a contiguous segment of the bucket and a private cursor
// per digit, so ranks need no atomics: __match_any groups the lanes of a step by digit.
__global__ void k_find_big(int nbuckets, const int *__restrict__ bstart, int cap, int *__restrict__ list,
                           int *__restrict__ soff, int *__restrict__ count_total)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbuckets) {
        const int nb = bstart[b + 1] - bstart[b];
        if (nb > cap) {
            const int k = atomicAdd(&count_total[0], 1);
            list[k] = b;
            soff[k] = atomicAdd(&count_total[1], nb);     // offset of this bucket's scratch slice
        }
    }
}

constexpr int BIG_THREADS = 1024;
constexpr int BIG_WARPS = BIG_THREADS / 32;
constexpr int BIG_UNROLL = 4;                  // 32-entry steps whose loads are in flight together
constexpr int BIG_MAX_WIDTH = 9;               // 512 digits x 32 warps x 4 B = 64 KB of cursors

template <bool VALUES>
__global__ void __launch_bounds__(BIG_THREADS)
k_bucket_big(int m, int log_rb, int nbuckets, int colbits, int width, int npasses,
             const int *__restrict__ big_list, const int *__restrict__ big_soff, int nbig,
             const int *__restrict__ bstart, int *ikey, double *ival, int *skey, double *sval,
             const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
             csi *Cp, csi *Ci, double *Cx)
{
    extern __shared__ __align__(16) int cursors[];             // [digit][warp]
    __shared__ int rowcnt[BK_RB_MAX + 1];
    __shared__ int rowstart[BK_RB_MAX + 1];
    __shared__ int warp_tot[BIG_WARPS];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned lt = lanemask_lt();
    const int rb = 1 << log_rb;
    const int nbins = 1 << width;
    // exclusive scan of one value per thread over the CTA
    auto block_excl = [&](int v) {
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
        __syncthreads();
        if (lane == 31) warp_tot[wid] = inc;
        __syncthreads();
        int before = inc - v;
        for (int w = 0; w < wid; w++) before += warp_tot[w];
        return before;
    };
    for (int idx = blockIdx.x; idx < nbig; idx += gridDim.x) {
        const int b = big_list[idx];
        const int base = bstart[b];
        const int nb = bstart[b + 1] - base;
        const int R0 = b << log_rb;
        const int nrows = min(rb, m - R0);
        // the packed word (local row, column) is the sort key as it stands
        int *keyA = ikey + base, *keyB = skey + big_soff[idx];
        double *valA = VALUES ? ival + base : nullptr, *valB = VALUES ? sval + big_soff[idx] : nullptr;
        const int colmask = (1 << colbits) - 1;
        // rows -> Cp (the reference's histogram + cs_cumsum restricted to this bucket)
        for (int k = tid; k <= rb; k += BIG_THREADS) rowcnt[k] = 0;
        __syncthreads();
        for (int e = tid; e < nb; e += BIG_THREADS) atomicAdd(&rowcnt[keyA[e] >> colbits], 1);
        __syncthreads();
        {
            const int c = tid < nrows ? rowcnt[tid] : 0;          // rb <= BK_RB_MAX == BIG_THREADS
            const int before = block_excl(c);
            if (tid <= nrows) rowstart[tid] = before;
            if (tid < nrows) Cp[R0 + tid] = base + before;
        }
        if (b == nbuckets - 1 && tid == 0) Cp[m] = base + nb;

        const int per = ((nb + BIG_WARPS - 1) / BIG_WARPS + 31) & ~31;     // segment of a warp
        const int seg_lo = min(nb, wid * per), seg_hi = min(nb, seg_lo + per);
        for (int pass = 0; pass < npasses; pass++) {
            const int *ksrc = (pass & 1) ? keyB : keyA;
            int *kdst = (pass & 1) ? keyA : keyB;
            const double *vsrc = (pass & 1) ? valB : valA;
            double *vdst = (pass & 1) ? valA : valB;
            const bool last = pass == npasses - 1;
            const int shift = pass * width;
            for (int k = tid; k < nbins * BIG_WARPS; k += BIG_THREADS) cursors[k] = 0;
            __syncthreads();
            // digit counts per warp segment
            for (int e0 = seg_lo; e0 < seg_hi; e0 += 32 * BIG_UNROLL) {
                int d[BIG_UNROLL];
#pragma unroll
                for (int u = 0; u < BIG_UNROLL; u++) {
                    const int e = e0 + u * 32 + lane;
                    d[u] = -1;
                    if (e < seg_hi) d[u] = (ksrc[e] >> shift) & (nbins - 1);
                }
#pragma unroll
                for (int u = 0; u < BIG_UNROLL; u++) {
                    // lanes past the end match nobody (digit nbins + lane); control flow stays uniform
                    const unsigned peers = __match_any_sync(0xffffffffu, d[u] >= 0 ? d[u] : nbins + lane);
                    if (d[u] >= 0 && (peers & lt) == 0) cursors[d[u] * BIG_WARPS + wid] += __popc(peers);
                    __syncwarp();
                }
            }
            __syncthreads();
            {   // exclusive scan of cursors in (digit, warp) order
                const int chunk = nbins * BIG_WARPS / BIG_THREADS;     // nbins >= 32 => chunk >= 1
                int sum = 0;
                for (int k = 0; k < chunk; k++) sum += cursors[tid * chunk + k];
                int before = block_excl(sum);
                for (int k = 0; k < chunk; k++) { const int c = cursors[tid * chunk + k]; cursors[tid * chunk + k] = before; before += c; }
            }
            __syncthreads();
            // stable scatter: every warp walks its segment in order
            for (int e0 = seg_lo; e0 < seg_hi; e0 += 32 * BIG_UNROLL) {
                int key[BIG_UNROLL], d[BIG_UNROLL];
                double val[BIG_UNROLL];
#pragma unroll
                for (int u = 0; u < BIG_UNROLL; u++) {
                    const int e = e0 + u * 32 + lane;
                    d[u] = -1;
                    key[u] = 0;
                    val[u] = 0.0;
                    if (e < seg_hi) {
                        key[u] = ksrc[e];
                        if (VALUES) val[u] = vsrc[e];
                        d[u] = (key[u] >> shift) & (nbins - 1);
                    }
                }
#pragma unroll
                for (int u = 0; u < BIG_UNROLL; u++) {
                    const bool valid = d[u] >= 0;
                    const unsigned peers = __match_any_sync(0xffffffffu, valid ? d[u] : nbins + lane);
                    const int leader = __ffs(peers) - 1;
                    int pos = 0;
                    if (valid && lane == leader) {
                        pos = cursors[d[u] * BIG_WARPS + wid];
                        cursors[d[u] * BIG_WARPS + wid] = pos + __popc(peers);
                    }
                    pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(peers & lt);
                    if (valid) {
                        if (!last) {
                            kdst[pos] = key[u];
                            if (VALUES) vdst[pos] = val[u];
                        } else {
                            Ci[base + pos] = key[u] & colmask;
                            if (VALUES) Cx[base + pos] = val[u];
                        }
                    }
                    __syncwarp();
                }
            }
            __syncthreads();
        }
        if (VALUES) {
            // duplicates of one (i,j) pair: their values go out in A's storage order
            for (int t = tid; t + 1 < nb; t += BIG_THREADS) {
                const int c = Ci[base + t];
                if (c != Ci[base + t + 1]) continue;
                const int rl = upper_row(rowstart, 0, nrows - 1, t);
                if (t + 1 >= rowstart[rl + 1]) continue;                      // next entry is another row's
                if (t > rowstart[rl] && Ci[base + t - 1] == c) continue;      // not the head of its group
                int g = 2;
                while (t + g < rowstart[rl + 1] && Ci[base + t + g] == c) g++;
                fix_tied_group(Ap, Ai, Ax, R0 + rl, c, Cx + base + t, g);
            }
        }
        __syncthreads();
    }
}

// ---- mirror path: square matrix, strictly increasing columns, symmetric pattern -----------------
// When every column of A is strictly increasing and (i, j) in A <=> (j, i) in A, the pattern of
// A' is the pattern of A (Cp = Ap, Ci = Ai), and the value of output position q -- row j = Ai[q]
// of output column i -- is A(i, j): the entry of source column j whose row is i.  Its position in
// column j is unique (no duplicates), so the transpose is ONE pass: stream Ai, look the mirror
// entry up, gather its value, write Ci / Cx coalesced.  24 B per entry of DRAM traffic (the
// algorithmic figure) instead of the 52 B of the two-hop bucket sort.  Source order inside an
// output column is ascending j (columns are visited in order, csparse.py:2308-2314), which is the
// order of the strictly increasing column i of A: bit-identical by construction.
// Nothing is assumed: the kernel checks every column for strict order and every entry for its
// mirror, and raises *fail otherwise (the caller then discards C and takes the general path).
// The lookup first tries position "same distance from the other end of the column" (exact for
// translation-invariant stencils away from the boundary: one 4-byte gather), then a binary search.
// SPEC: the value at the guessed position is requested together with the probe (one exposed memory phase
// less per tile) instead of after every probe of the tile has been checked.  Measured: lap2d 4096^2
// 0.81 -> 0.77 ms, but st27 128^3 0.71 -> 0.77 ms (twice the gathers in flight on lines the probes also
// want), so the host turns it on for short columns only (<= 8 entries on average).
template <bool VALUES, bool SPEC>
__global__ void __launch_bounds__(TR_THREADS, 4)
k_mirror(const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
         int n, long long nnz, int ntiles, const int *__restrict__ tile_col,
         csi *__restrict__ Ci, double *__restrict__ Cx, int *fail)
{
    __shared__ int sAp[PT_SMEM_COLS];
    const int lane = threadIdx.x & 31;
    int round = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, round++) {
        // block-uniform early exit once anyone found an asymmetry (looked at every 8th tile: the flag
        // is an L2 round trip); also the barrier that lets sAp be restaged
        if ((round & 7) == 0) {
            if (__syncthreads_or(*reinterpret_cast<volatile int *>(fail))) return;
        } else {
            __syncthreads();
        }
        const long long p_begin = (long long)t * PT_TILE;
        const long long p_end = min(nnz, p_begin + PT_TILE);
        const int j_first = tile_col[t];
        const int j_last = tile_col[t + 1];
        const int ncols = j_last - j_first + 1;
        const bool staged = ncols + 1 <= PT_SMEM_COLS;
        if (staged)
            for (int k = threadIdx.x; k <= ncols; k += TR_THREADS) sAp[k] = Ap[j_first + k];
        __syncthreads();

        int rows[PT_EPT];       // row index of the entry = the source column of its mirror
        int off[PT_EPT];        // distance from the start of the entry's own column
        int col[PT_EPT];        // the entry's own column = the row of its mirror
        int src[PT_EPT];        // position of the mirror entry
        int cntk[PT_EPT / 4];
        bool bad = false;
#pragma unroll
        for (int k = 0; k < PT_EPT / 4; k++) {
            const long long p = p_begin + (long long)(k * TR_THREADS + threadIdx.x) * 4;
            cntk[k] = p < p_end ? (int)min((long long)4, p_end - p) : 0;
            if (cntk[k] == 4) {
                const int4 r = ldg_stream(reinterpret_cast<const int4 *>(Ai + p));
                rows[4 * k] = r.x; rows[4 * k + 1] = r.y; rows[4 * k + 2] = r.z; rows[4 * k + 3] = r.w;
            } else {
#pragma unroll
                for (int e = 0; e < 4; e++) rows[4 * k + e] = e < cntk[k] ? Ai[p + e] : -1;
            }
        }
        // own column of every entry; strict order inside the column
#pragma unroll
        for (int k = 0; k < PT_EPT / 4; k++) {
            const int p = (int)(p_begin + (long long)(k * TR_THREADS + threadIdx.x) * 4);
            int prev = __shfl_up_sync(0xffffffffu, rows[4 * k + 3], 1);      // entry p - 1 (full group)
            if (cntk[k] > 0) {
                if (lane == 0 && p > 0) prev = Ai[p - 1];
                int j;   // largest j with Ap[j] <= p
                if (staged) j = upper_row(sAp, 0, ncols - 1, p);
                else        j = upper_row(Ap, j_first, j_last, p) - j_first;
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (e < cntk[k]) {
                        const int pe = p + e;
                        if (staged) { while (sAp[j + 1] <= pe) j++; }
                        else        { while (Ap[j_first + j + 1] <= pe) j++; }
                        const int o = pe - (staged ? sAp[j] : Ap[j_first + j]);
                        off[4 * k + e] = o;
                        col[4 * k + e] = j_first + j;
                        const int r = rows[4 * k + e];
                        if (o > 0 && prev >= r) bad = true;                    // not strictly increasing
                        if ((unsigned)r >= (unsigned)n) bad = true;
                        prev = r;
                    }
                }
            }
        }
        unsigned miss = 0;      // entries whose guessed position does not hold the mirror
        double v[PT_EPT];       // the value at the guessed position travels with the probe (one memory phase less per tile)
        if (!bad) {
            // guess: as far from the end of the mirror column as this entry is from its own column's
            // start.  The column bounds are gathers (16 loads in flight), then one 4-byte probe each.
#pragma unroll
            for (int k = 0; k < PT_EPT; k++) {
                src[k] = -1;
                if (k % 4 < cntk[k / 4]) {
                    const unsigned rel = (unsigned)(rows[k] - j_first);     // mirror column inside the tile's own range:
                    int ca, cb;                                              // its bounds are already in shared memory
                    if (staged && rel < (unsigned)ncols) { ca = sAp[rel]; cb = sAp[rel + 1]; }
                    else { ca = Ap[rows[k]]; cb = Ap[rows[k] + 1]; }
                    const int g = cb - 1 - off[k];
                    src[k] = g >= ca ? g : -1;
                }
            }
#pragma unroll
            for (int k = 0; k < PT_EPT; k++) {
                if (k % 4 < cntk[k / 4]) {
                    const int got = src[k] >= 0 ? Ai[src[k]] : -1;
                    if (VALUES && SPEC) v[k] = src[k] >= 0 ? Ax[src[k]] : 0.0;
                    if (got != col[k]) miss |= 1u << k;
                }
            }
        }
        if (miss) {             // boundary columns, irregular patterns: binary search in the mirror column
#pragma unroll
            for (int k = 0; k < PT_EPT; k++) {
                if (miss >> k & 1) {
                    const int cb = Ap[rows[k] + 1];
                    int lo = Ap[rows[k]], hi = cb;                              // first position with Ai >= col
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (Ai[mid] < col[k]) lo = mid + 1; else hi = mid;
                    }
                    if (lo < cb && Ai[lo] == col[k]) { src[k] = lo; if (VALUES && SPEC) v[k] = Ax[lo]; } else bad = true;
                }
            }
        }
        if (bad) {
            *fail = 1;
        } else {
            if (VALUES && !SPEC) {
#pragma unroll
                for (int k = 0; k < PT_EPT; k++) if (k % 4 < cntk[k / 4]) v[k] = Ax[src[k]];
            }
#pragma unroll
            for (int k = 0; k < PT_EPT / 4; k++) {
                const long long p = p_begin + (long long)(k * TR_THREADS + threadIdx.x) * 4;
                if (cntk[k] == 4) {
                    *reinterpret_cast<int4 *>(Ci + p) = make_int4(rows[4 * k], rows[4 * k + 1], rows[4 * k + 2], rows[4 * k + 3]);
                    if (VALUES) stg4(Cx + p, v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; e++)
                        if (e < cntk[k]) { Ci[p + e] = rows[4 * k + e]; if (VALUES) Cx[p + e] = v[4 * k + e]; }
                }
            }
        }
    }
}

// fused launch: a bucket without entries is completed by nobody; its rows' column pointers are set here
__global__ void k_empty_buckets(int m, int log_rb, int nbuckets, const int *__restrict__ bstart, csi *__restrict__ Cp)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    const int base = bstart[b];
    if (bstart[b + 1] != base) return;
    const int R0 = b << log_rb, nrows = min(1 << log_rb, m - R0);
    for (int r = 0; r < nrows; r++) Cp[R0 + r] = base;
    if (b == nbuckets - 1) Cp[m] = base;
}

// ---- host side -------------------------------------------------------------------
thread_local int t_last_path = 0;   // 1 mirror, 2 bucket sort, 3 radix sort, 0 trivial
int transpose_last_path() { return t_last_path; }

int transpose_impl(const csb200_mat *A, bool values, csb200_mat **out)
{
    const csi m = A->m, n = A->n;
    const long long nnz = A->nnz;
    const bool has_x = values && A->x != nullptr;
    const size_t cap = (size_t)(nnz > 0 ? nnz : 1) + MAT_PAD;

    csb200_mat *C = new csb200_mat();
    C->m = n; C->n = m; C->nnz = nnz; C->device = A->device;
    int st = dev_alloc(&C->p, (size_t)m + 1 + MAT_PAD);
    if (st == CSB200_OK) st = dev_alloc(&C->i, cap);
    if (st == CSB200_OK && has_x) st = dev_alloc(&C->x, cap);
    auto fail = [&](int s) { csb200_mat_free(C); return s; };
    if (st != CSB200_OK) return fail(st);
    cudaStream_t s = stream();
#define TR_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
        set_error(CSB200_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); return fail(CSB200_ERR_CUDA); } } while (0)
#define TR_LAUNCHED() do { g_launches.fetch_add(1, std::memory_order_relaxed); TR_CUDA(cudaGetLastError()); } while (0)
    t_last_path = 0;
    if (nnz == 0 || m == 0) {   // cs_spalloc leaves one zero slot (csparse.py:2401); Cp is all zero
        TR_CUDA(cudaMemsetAsync(C->p, 0, ((size_t)m + 1) * sizeof(csi), s));
        TR_CUDA(cudaMemsetAsync(C->i, 0, sizeof(csi), s));
        if (has_x) TR_CUDA(cudaMemsetAsync(C->x, 0, sizeof(double), s));
        *out = C;
        return CSB200_OK;
    }

    // rows per bucket: the largest power of two whose average bucket fills <= 85 % of the staging area
    const double avg = (double)nnz / m;
    // short rows: one warp per small bucket; longer rows: one CTA per larger bucket
    const bool warp_path = avg <= 8.0;       // (one warp per bucket with 16 entries per lane on the 27-point stencil: 1.63 ms against 1.50)
    const int bcap = warp_path ? WB_CAP : BK_CAP;
    int log_rb = 2;
    while (log_rb < (warp_path ? 7 : 10) && (double)(2 << log_rb) * avg <= 0.85 * bcap) log_rb++;
    const int nbuckets = (int)(((long long)m + (1 << log_rb) - 1) >> log_rb);
    const int ntiles = ceil_div(nnz, PT_TILE);
    // the bucket path moves one packed word (row inside the bucket, source column) per entry
    int colbits = 1;
    while (colbits < 31 && (1LL << colbits) < (long long)n) colbits++;
    const bool packable = colbits + log_rb <= 31;

    // temporaries of the bucket path: one packed word + one value per entry, bucket tables, tile table
    arena_hint((size_t)nnz * 12 + (size_t)nbuckets * 24 + (size_t)ntiles * 4 + (1 << 16));
    DevBuf<int> bstart, bfill, tile_col;
    DevBuf<long long> total;
    if ((st = bstart.alloc((size_t)nbuckets + 1)) || (st = bfill.alloc((size_t)nbuckets + 2)) ||
        (st = tile_col.alloc((size_t)ntiles + 1)) || (st = total.alloc(1)))
        return fail(st);
    auto radix_path = [&]() {
        int st2 = stable_sort_by_key(nnz, m, A->i, nullptr, A->p, n, has_x ? A->x : nullptr, C->p, C->i, C->x);
        if (st2 != CSB200_OK) return fail(st2);
        t_last_path = 3;
        *out = C;
        return (int)CSB200_OK;
    };
    // one-pass mirror path: tried on square matrices until the handle is known not to qualify; the
    // kernel verifies what it relies on (strictly increasing columns, symmetric pattern) and backs out
    if (tls().force_transpose == 0 && m == n && A->mirror != 0) {
        DevBuf<int> flag;
        if ((st = flag.alloc(1)) != CSB200_OK) return fail(st);
        TR_CUDA(cudaMemsetAsync(flag.ptr, 0, sizeof(int), s));
        k_tile_cols<<<ceil_div(ntiles + 1, 256), 256, 0, s>>>(A->p, n, nnz, ntiles, tile_col.ptr);
        TR_LAUNCHED();
        // resident CTAs striding over the tiles
        auto launch = [&](auto kern) -> int {
            int per_sm = 4;
            cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, TR_THREADS, 0);
            if (e != cudaSuccess) return set_error(CSB200_ERR_CUDA, "occupancy query: %s", cudaGetErrorString(e));
            const int grid = min(ntiles, sm_count() * max(per_sm, 1));
            kern<<<grid, TR_THREADS, 0, s>>>(A->p, A->i, has_x ? A->x : nullptr, n, nnz, ntiles, tile_col.ptr, C->i,
                                             has_x ? C->x : nullptr, flag.ptr);
            return CSB200_OK;
        };
        st = !has_x ? launch(k_mirror<false, false>) : (nnz <= 8LL * n ? launch(k_mirror<true, true>) : launch(k_mirror<true, false>));
        if (st != CSB200_OK) return fail(st);
        TR_LAUNCHED();
        TR_CUDA(cudaMemcpyAsync(C->p, A->p, ((size_t)n + 1) * sizeof(csi), cudaMemcpyDeviceToDevice, s));
        int h_flag = 1;
        TR_CUDA(cudaMemcpyAsync(&h_flag, flag.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
        TR_CUDA(cudaStreamSynchronize(s));
        const_cast<csb200_mat *>(A)->mirror = h_flag ? 0 : 1;
        if (!h_flag) {
            const_cast<csb200_mat *>(A)->canon = 1;
            C->canon = 1;
            C->mirror = 1;
            t_last_path = 1;
            *out = C;
            return CSB200_OK;
        }
    }
    if (tls().force_transpose == 1 || !packable || A->wide_rows == 1) return radix_path();      // wide_rows: found by an earlier call's histogram
    TR_CUDA(cudaMemsetAsync(bfill.ptr, 0, ((size_t)nbuckets + 2) * sizeof(int), s));
    const int nht = ceil_div(nnz, TR_TILE);
    std::vector<int> h_range((size_t)2 * nht);
    {
        int *wide = bfill.ptr + nbuckets + 1, h_wide = 0;
        DevBuf<int> tile_range;
        if ((st = tile_range.alloc((size_t)2 * nht)) != CSB200_OK) return fail(st);
        k_bucket_hist<<<nht, TR_THREADS, 0, s>>>(A->i, nnz, log_rb, bfill.ptr, wide, tile_range.ptr);
        TR_LAUNCHED();
        TR_CUDA(cudaMemcpyAsync(&h_wide, wide, sizeof(int), cudaMemcpyDeviceToHost, s));
        TR_CUDA(cudaMemcpyAsync(h_range.data(), tile_range.ptr, h_range.size() * sizeof(int), cudaMemcpyDeviceToHost, s));
        TR_CUDA(cudaStreamSynchronize(s));
        const_cast<csb200_mat *>(A)->wide_rows = h_wide > 0 ? 1 : 0;      // a fact about A, remembered like `mirror`
        if (h_wide > 0) return radix_path();
    }
    // bstart = exclusive scan of the counts; bfill <- bstart (the fill cursors)
    if ((st = launch_excl_scan(bstart.ptr, bfill.ptr, nbuckets, total.ptr, nullptr)) != CSB200_OK) return fail(st);
    // oversized buckets (power-law rows): a few are sorted one CTA each (k_bucket_big); when they
    // hold a sizeable share of the matrix the whole transpose goes through the stable radix sort
    DevBuf<int> big_list, big_soff, big_ct;
    int h_ct[2] = {0, 0};
    if (nnz > bcap) {
        if ((st = big_list.alloc((size_t)nbuckets)) || (st = big_soff.alloc((size_t)nbuckets)) ||
            (st = big_ct.alloc(2)))
            return fail(st);
        TR_CUDA(cudaMemsetAsync(big_ct.ptr, 0, 2 * sizeof(int), s));
        k_find_big<<<ceil_div(nbuckets, 256), 256, 0, s>>>(nbuckets, bstart.ptr, bcap, big_list.ptr, big_soff.ptr, big_ct.ptr);
        TR_LAUNCHED();
        TR_CUDA(cudaMemcpyAsync(h_ct, big_ct.ptr, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        TR_CUDA(cudaStreamSynchronize(s));
        if ((long long)h_ct[1] * 8 > nnz) return radix_path();
    }
    // intermediate: one packed word (row inside the bucket, source column) per entry + its value
    DevBuf<int> ikey;
    DevBuf<double> ival;
    if ((st = ikey.alloc((size_t)nnz + 8)) != CSB200_OK) return fail(st);
    if (has_x && (st = ival.alloc((size_t)nnz + 8)) != CSB200_OK) return fail(st);
    k_tile_cols<<<ceil_div(ntiles + 1, 256), 256, 0, s>>>(A->p, n, nnz, ntiles, tile_col.ptr);
    TR_LAUNCHED();
    // Slabs (opt-in, csb200_transpose_force_path(3)): a banded matrix completes its buckets in order
    // -- once the entries up to some tile are partitioned, no later tile feeds the buckets below the
    // smallest bucket those later tiles touch.  The partition can therefore be cut into slabs of a
    // few million entries, the buckets a slab completes being sorted at once while their share of
    // the intermediate is still in the 126 MB L2.  MEASURED (B200, profiles/r2_notes.md): the ~40
    // short launches cost more in ramp-up and tail than the second hop saves -- lap2d 4096^2 1.77 ms
    // against 1.34 ms for two whole passes, st27 128^3 1.88 against 1.50 -- so the default is one slab.
    struct Slab { int tile_end, bucket_end; };
    std::vector<Slab> slabs;
    {
        constexpr long long SLAB_LAG_BYTES = 40LL << 20;
        const long long slab_entries = std::min<long long>(4LL << 20, std::max<long long>(1LL << 20, nnz / 16));
        const int step = (int)(slab_entries / TR_TILE);
        bool ok = tls().force_transpose == 3 && h_ct[0] == 0 && nht >= 4 * step;
        if (ok) {
            std::vector<int> sufmin((size_t)nht + 1);
            sufmin[nht] = nbuckets;
            for (int t = nht - 1; t >= 0; t--) sufmin[t] = std::min(sufmin[t + 1], h_range[2 * t]);
            const double bytes_per_bucket = (has_x ? 12.0 : 4.0) * (double)nnz / nbuckets;
            int premax = -1, t = 0;
            for (int e = step; ok; e += step) {
                if (e > nht) e = nht;
                for (; t < e; t++) premax = std::max(premax, h_range[2 * t + 1]);
                const int frontier = e == nht ? nbuckets : sufmin[e];
                if ((double)(premax + 1 - frontier) * bytes_per_bucket > (double)SLAB_LAG_BYTES) ok = false;
                slabs.push_back({std::min(2 * e, ntiles), frontier});
                if (e == nht) break;
            }
        }
        if (!ok) { slabs.clear(); slabs.push_back({ntiles, nbuckets}); }
    }
    TR_CUDA(cudaFuncSetAttribute(k_bucket_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM));
    TR_CUDA(cudaFuncSetAttribute(k_bucket_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BK_SMEM));
    // opt-in (force_path 4), warp-per-bucket matrices: partition and sort in ONE persistent launch
    // (k_partition_sort); the second hop reads L2.  MEASURED: 1.94 ms against 1.35 ms for the two whole
    // passes on lap2d 4096^2 -- three CTAs per SM at 80 registers + 72 KB, and a CTA sorts its ~13
    // completed buckets with eight warps only after its tile is partitioned, where the two kernels
    // each run at their own best occupancy.  Not the default.
    const bool fused = warp_path && h_ct[0] == 0 && slabs.size() == 1 && tls().force_transpose == 4;
    if (fused) {
        DevBuf<int> arrived;
        DevBuf<unsigned> ticket;
        if ((st = arrived.alloc((size_t)nbuckets + 1)) || (st = ticket.alloc(1))) return fail(st);
        TR_CUDA(cudaMemsetAsync(arrived.ptr, 0, ((size_t)nbuckets + 1) * sizeof(int), s));
        TR_CUDA(cudaMemsetAsync(ticket.ptr, 0, sizeof(unsigned), s));
        constexpr int smem = WB_WARPS * WB_WARP_BYTES;
        TR_CUDA(cudaFuncSetAttribute(k_partition_sort<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        TR_CUDA(cudaFuncSetAttribute(k_partition_sort<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (has_x) {
            const int grid = min(ntiles, resident_grid(k_partition_sort<true>, TR_THREADS, smem));
            k_partition_sort<true><<<grid, TR_THREADS, smem, s>>>(A->p, A->i, A->x, nnz, ntiles, tile_col.ptr, m, log_rb, colbits, nbuckets,
                                                                  bfill.ptr, bstart.ptr, arrived.ptr, ticket.ptr, ikey.ptr, ival.ptr, C->p, C->i, C->x);
        } else {
            const int grid = min(ntiles, resident_grid(k_partition_sort<false>, TR_THREADS, smem));
            k_partition_sort<false><<<grid, TR_THREADS, smem, s>>>(A->p, A->i, nullptr, nnz, ntiles, tile_col.ptr, m, log_rb, colbits, nbuckets,
                                                                   bfill.ptr, bstart.ptr, arrived.ptr, ticket.ptr, ikey.ptr, nullptr, C->p, C->i, nullptr);
        }
        TR_LAUNCHED();
        k_empty_buckets<<<ceil_div(nbuckets, 256), 256, 0, s>>>(m, log_rb, nbuckets, bstart.ptr, C->p);
        TR_LAUNCHED();
        slabs.clear();
    }
    int tile0 = 0, b0 = 0;
    for (const Slab &sl : slabs) {
        const int nt = sl.tile_end - tile0;
        if (nt > 0) {
            if (has_x) k_partition<true><<<nt, TR_THREADS, 0, s>>>(A->p, A->i, A->x, nnz, tile_col.ptr, log_rb, colbits, bfill.ptr, ikey.ptr, ival.ptr, tile0);
            else       k_partition<false><<<nt, TR_THREADS, 0, s>>>(A->p, A->i, nullptr, nnz, tile_col.ptr, log_rb, colbits, bfill.ptr, ikey.ptr, nullptr, tile0);
            TR_LAUNCHED();
        }
        tile0 = sl.tile_end;
        const int b1 = sl.bucket_end;
        if (b1 > b0 && warp_path) {
            constexpr int smem = WB_WARPS * WB_WARP_BYTES;
            const int grid = min(ceil_div(b1 - b0, WB_WARPS), sm_count() * 4);
            if (has_x) k_bucket_sort_warp<true><<<grid, WB_WARPS * 32, smem, s>>>(m, log_rb, colbits, nbuckets, bstart.ptr, ikey.ptr, ival.ptr, A->p, A->i, A->x, C->p, C->i, C->x, b0, b1);
            else       k_bucket_sort_warp<false><<<grid, WB_WARPS * 32, smem, s>>>(m, log_rb, colbits, nbuckets, bstart.ptr, ikey.ptr, nullptr, A->p, A->i, nullptr, C->p, C->i, nullptr, b0, b1);
            TR_LAUNCHED();
        } else if (b1 > b0) {
            if (has_x) k_bucket_sort<true><<<b1 - b0, BK_THREADS, BK_SMEM, s>>>(m, log_rb, colbits, nbuckets, bstart.ptr, ikey.ptr, ival.ptr, A->p, A->i, A->x, C->p, C->i, C->x, b0);
            else       k_bucket_sort<false><<<b1 - b0, BK_THREADS, BK_SMEM, s>>>(m, log_rb, colbits, nbuckets, bstart.ptr, ikey.ptr, nullptr, A->p, A->i, nullptr, C->p, C->i, nullptr, b0);
            TR_LAUNCHED();
        }
        b0 = std::max(b0, b1);
    }
    {
        if (h_ct[0] > 0) {
            // key = the packed word (local row, source column): colbits + log_rb bits, in an odd number of passes
            const int bits = colbits + log_rb;
            const int npasses = bits <= BIG_MAX_WIDTH ? 1 : bits <= 3 * BIG_MAX_WIDTH ? 3 : 5;
            int width = (bits + npasses - 1) / npasses;
            if (width < 5) width = 5;                    // the scan wants one cursor per thread at least
            DevBuf<int> skey;
            DevBuf<double> sval;
            if ((st = skey.alloc((size_t)h_ct[1])) != CSB200_OK) return fail(st);
            if (has_x && (st = sval.alloc((size_t)h_ct[1])) != CSB200_OK) return fail(st);
            const int smem = (int)((sizeof(int) * BIG_WARPS) << width);
            const int grid = min(h_ct[0], sm_count());
            TR_CUDA(cudaFuncSetAttribute(k_bucket_big<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            TR_CUDA(cudaFuncSetAttribute(k_bucket_big<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            if (has_x) k_bucket_big<true><<<grid, BIG_THREADS, smem, s>>>(m, log_rb, nbuckets, colbits, width, npasses, big_list.ptr, big_soff.ptr, h_ct[0], bstart.ptr, ikey.ptr, ival.ptr, skey.ptr, sval.ptr, A->p, A->i, A->x, C->p, C->i, C->x);
            else       k_bucket_big<false><<<grid, BIG_THREADS, smem, s>>>(m, log_rb, nbuckets, colbits, width, npasses, big_list.ptr, big_soff.ptr, h_ct[0], bstart.ptr, ikey.ptr, nullptr, skey.ptr, nullptr, A->p, A->i, nullptr, C->p, C->i, nullptr);
            TR_LAUNCHED();
        }
    }
#undef TR_CUDA
#undef TR_LAUNCHED
    t_last_path = 2;
    *out = C;
    return CSB200_OK;
}

}  // namespace csb
