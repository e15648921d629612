// transpose.cu -- cs_transpose (csparse.py:2292-2315): C = A' as a counting sort
// by row index.  The reference is a sequential stable scatter; the result must be
// bit-identical (p, i and x), so the order inside every output column has to be
// the source order.
//
// Kernels (all HBM-bound integer / byte movement, no tensor cores):
//   k_hist        row histogram           w[Ai[p]]++               (:2305-2306)
//   excl_scan     Cp = cumsum(w), w = Cp  (scan.cu)                (:2307)
//   k_tile_cols   column of the first entry of every 4096-entry tile
//   k_scatter     q = w[Ai[p]]++ ; Ci[q] = j ; Cx[q] = Ax[p]       (:2308-2314)
//   k_fix_*       restore source order inside each output column
//
// The scatter hands out slots with atomics, so entries of one output column may
// land permuted.  Entries of an output column come from distinct source columns
// (or are duplicates of one (i,j) pair), hence source order == ascending j with
// ties in storage order: the fix kernels check every output column, sort the few
// that are out of order by j, and re-read tied groups from the source column.
// Correctness never depends on how the atomics were ordered; only speed does.
#include "common.cuh"

namespace csb {

constexpr int TR_THREADS = 256;
constexpr int TR_TILE = 4096;                 // entries per scatter CTA
constexpr int TR_SMEM_COLS = 6144;            // column pointers staged per tile
constexpr int FIX_SHORT = 32;                 // rows up to this length: one thread
constexpr int FIX_MID = 1024;                 // up to this: one warp; longer: one CTA

// ---- histogram ---------------------------------------------------------------
__global__ void __launch_bounds__(TR_THREADS)
k_hist(const csi *__restrict__ Ai, long long nnz, int *__restrict__ w)
{
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    for (long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; p < nnz; p += stride) {
        if (p + 3 < nnz) {
            const int4 r = ldg_stream(reinterpret_cast<const int4 *>(Ai + p));
            atomicAdd(&w[r.x], 1);
            atomicAdd(&w[r.y], 1);
            atomicAdd(&w[r.z], 1);
            atomicAdd(&w[r.w], 1);
        } else {
            for (long long q = p; q < nnz; q++) atomicAdd(&w[Ai[q]], 1);
        }
    }
}

// ---- tile -> first column ------------------------------------------------------
__global__ void k_tile_cols(const csi *__restrict__ Ap, int n, long long nnz, int ntiles,
                            int *__restrict__ tile_col)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > ntiles) return;
    if (t == ntiles) { tile_col[t] = n - 1; return; }
    const long long p0 = (long long)t * TR_TILE;
    tile_col[t] = upper_row(Ap, 0, n, (int)p0);   // largest j with Ap[j] <= p0
}

// ---- scatter -------------------------------------------------------------------
template <bool VALUES>
__global__ void __launch_bounds__(TR_THREADS)
k_scatter(const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
          long long nnz, const int *__restrict__ tile_col, int *__restrict__ w,
          csi *__restrict__ Ci, double *__restrict__ Cx)
{
    __shared__ int sAp[TR_SMEM_COLS];
    const int t = blockIdx.x;
    const long long p_begin = (long long)t * TR_TILE;
    const long long p_end = min(nnz, p_begin + TR_TILE);
    const int j_first = tile_col[t];
    const int j_last = tile_col[t + 1];           // >= column of the last entry of this tile
    const int ncols = j_last - j_first + 1;
    const bool staged = ncols + 1 <= TR_SMEM_COLS;
    if (staged)
        for (int k = threadIdx.x; k <= ncols; k += TR_THREADS) sAp[k] = Ap[j_first + k];
    __syncthreads();

    for (long long p = p_begin + threadIdx.x * 4; p < p_end; p += TR_THREADS * 4) {
        int rows[4];
        double vals[4];
        const int cnt = (int)min((long long)4, p_end - p);
        if (cnt == 4) {
            const int4 r = ldg_stream(reinterpret_cast<const int4 *>(Ai + p));
            rows[0] = r.x; rows[1] = r.y; rows[2] = r.z; rows[3] = r.w;
            if (VALUES) {
                const double2 a = ldg_stream(reinterpret_cast<const double2 *>(Ax + p));
                const double2 b = ldg_stream(reinterpret_cast<const double2 *>(Ax + p + 2));
                vals[0] = a.x; vals[1] = a.y; vals[2] = b.x; vals[3] = b.y;
            }
        } else {
            for (int e = 0; e < cnt; e++) {
                rows[e] = Ai[p + e];
                if (VALUES) vals[e] = Ax[p + e];
            }
        }
        int j;   // column of entry p: largest j with Ap[j] <= p
        if (staged) j = upper_row(sAp, 0, ncols - 1, (int)p);
        else        j = upper_row(Ap, j_first, j_last, (int)p) - j_first;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            if (e < cnt) {
                const int pe = (int)p + e;
                if (staged) { while (sAp[j + 1] <= pe) j++; }
                else        { while (Ap[j_first + j + 1] <= pe) j++; }
                const int q = atomicAdd(&w[rows[e]], 1);
                Ci[q] = j_first + j;
                if (VALUES) Cx[q] = vals[e];
            }
        }
    }
}

// ---- order repair ----------------------------------------------------------------
// After sorting an output column by j, runs of equal j are duplicates of one
// (row r, column j) entry of A; their values must appear in A's storage order.
__device__ void fix_tied_group(const csi *Ap, const csi *Ai, const double *Ax,
                               int r, int j, double *cx, int g)
{
    int k = 0;
    for (int p = Ap[j]; p < Ap[j + 1] && k < g; p++)
        if (Ai[p] == r) cx[k++] = Ax[p];
}

template <bool VALUES>
__global__ void __launch_bounds__(256)
k_fix_short(int m, const csi *__restrict__ Cp, csi *Ci, double *Cx,
            const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
            int *mid_list, int *big_list, int *counts)
{
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    const int b = Cp[r], len = Cp[r + 1] - b;
    if (len < 2) return;
    if (len > FIX_SHORT) {
        if (len <= FIX_MID) mid_list[atomicAdd(&counts[0], 1)] = r;
        else                big_list[atomicAdd(&counts[1], 1)] = r;
        return;
    }
    bool sorted = true;
    int prev = Ci[b];
    for (int k = 1; k < len; k++) {
        const int cur = Ci[b + k];
        sorted &= cur > prev;
        prev = cur;
    }
    if (sorted) return;
    // rare: insertion sort of (j, x) pairs by j
    int key[FIX_SHORT];
    double val[FIX_SHORT];
    for (int k = 0; k < len; k++) {
        key[k] = Ci[b + k];
        if (VALUES) val[k] = Cx[b + k];
    }
    for (int a = 1; a < len; a++) {
        const int kj = key[a];
        const double kv = VALUES ? val[a] : 0.0;
        int c = a - 1;
        while (c >= 0 && key[c] > kj) {
            key[c + 1] = key[c];
            if (VALUES) val[c + 1] = val[c];
            c--;
        }
        key[c + 1] = kj;
        if (VALUES) val[c + 1] = kv;
    }
    for (int k = 0; k < len; k++) {
        Ci[b + k] = key[k];
        if (VALUES) Cx[b + k] = val[k];
    }
    if (VALUES) {
        for (int k = 0; k + 1 < len;) {
            int g = 1;
            while (k + g < len && key[k + g] == key[k]) g++;
            if (g > 1) fix_tied_group(Ap, Ai, Ax, r, key[k], Cx + b + k, g);
            k += g;
        }
    }
}

// Cooperative in-place sort of one output column by a group of G threads
// (G = 32: a warp, G = blockDim: a CTA).  Normalised bitonic network: every
// comparator moves the smaller key to the lower index, so virtual +inf padding
// above `len` never moves and any length works.
template <int G, bool VALUES>
__device__ void group_sort_row(csi *ci, double *cx, int len, int tid)
{
    int P = 2;
    while (P < len) P <<= 1;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (P >> 1); t += G) {
                // t-th comparator of this stage
                const int lo = ((t / j) * (j << 1)) + (t % j);
                const int hi = (j == (k >> 1)) ? (lo ^ (k - 1)) : (lo + j);
                // for the mirror stage lo^(k-1) > lo always holds (lo's bit j is 0)
                if (hi < len) {
                    const int a = ci[lo], b = ci[hi];
                    if (a > b) {
                        ci[lo] = b; ci[hi] = a;
                        if (VALUES) { const double xa = cx[lo]; cx[lo] = cx[hi]; cx[hi] = xa; }
                    }
                }
            }
            if (G == 32) __syncwarp(); else __syncthreads();
        }
    }
}

template <int G, bool VALUES>
__device__ void group_fix_row(int r, const csi *Cp, csi *Ci, double *Cx,
                              const csi *Ap, const csi *Ai, const double *Ax, int tid, int *flag)
{
    const int b = Cp[r], len = Cp[r + 1] - b;
    csi *ci = Ci + b;
    double *cx = VALUES ? Cx + b : nullptr;
    // cooperative order check
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
    if (!bad) return;
    group_sort_row<G, VALUES>(ci, cx, len, tid);
    if (VALUES) {
        for (int t = tid; t + 1 < len; t += G) {
            if (ci[t] == ci[t + 1] && (t == 0 || ci[t - 1] != ci[t])) {
                int g = 2;
                while (t + g < len && ci[t + g] == ci[t]) g++;
                fix_tied_group(Ap, Ai, Ax, r, ci[t], cx + t, g);
            }
        }
    }
}

template <bool VALUES>
__global__ void __launch_bounds__(128)
k_fix_mid(const int *list, const int *counts, const csi *Cp, csi *Ci, double *Cx,
          const csi *Ap, const csi *Ai, const double *Ax)
{
    const int nrows = counts[0];
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < nrows; k += warps)
        group_fix_row<32, VALUES>(list[k], Cp, Ci, Cx, Ap, Ai, Ax, threadIdx.x & 31, nullptr);
}

template <bool VALUES>
__global__ void __launch_bounds__(512)
k_fix_big(const int *list, const int *counts, const csi *Cp, csi *Ci, double *Cx,
          const csi *Ap, const csi *Ai, const double *Ax)
{
    __shared__ int flag;
    const int nrows = counts[1];
    for (int k = blockIdx.x; k < nrows; k += gridDim.x)
        group_fix_row<512, VALUES>(list[k], Cp, Ci, Cx, Ap, Ai, Ax, threadIdx.x, &flag);
}

// ---- host side -------------------------------------------------------------------
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

    DevBuf<int> w, tile_col, lists, counts;
    DevBuf<long long> total;
    if ((st = w.alloc((size_t)(m > 0 ? m : 1))) != CSB200_OK) return fail(st);
    if ((st = total.alloc(1)) != CSB200_OK) return fail(st);
    cudaStream_t s = stream();
#define TR_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
        set_error(CSB200_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); return fail(CSB200_ERR_CUDA); } } while (0)
#define TR_LAUNCHED() do { g_launches.fetch_add(1, std::memory_order_relaxed); TR_CUDA(cudaGetLastError()); } while (0)
    TR_CUDA(cudaMemsetAsync(w.ptr, 0, (size_t)(m > 0 ? m : 1) * sizeof(int), s));
    if (nnz == 0) {   // cs_spalloc leaves one zero slot (csparse.py:2401)
        TR_CUDA(cudaMemsetAsync(C->i, 0, sizeof(csi), s));
        if (has_x) TR_CUDA(cudaMemsetAsync(C->x, 0, sizeof(double), s));
    }
    if (nnz > 0) {
        const int blocks = (int)min((long long)ceil_div(nnz, TR_THREADS * 4 * 4), (long long)148 * 64);
        k_hist<<<blocks, TR_THREADS, 0, s>>>(A->i, nnz, w.ptr);
        TR_LAUNCHED();
    }
    if ((st = launch_excl_scan(C->p, w.ptr, m, total.ptr, nullptr)) != CSB200_OK) return fail(st);
    if (nnz > 0) {
        const int ntiles = ceil_div(nnz, TR_TILE);
        if ((st = tile_col.alloc((size_t)ntiles + 1)) != CSB200_OK) return fail(st);
        k_tile_cols<<<ceil_div(ntiles + 1, 256), 256, 0, s>>>(A->p, n, nnz, ntiles, tile_col.ptr);
        TR_LAUNCHED();
        if (has_x)
            k_scatter<true><<<ntiles, TR_THREADS, 0, s>>>(A->p, A->i, A->x, nnz, tile_col.ptr, w.ptr, C->i, C->x);
        else
            k_scatter<false><<<ntiles, TR_THREADS, 0, s>>>(A->p, A->i, nullptr, nnz, tile_col.ptr, w.ptr, C->i, nullptr);
        TR_LAUNCHED();

        // order repair
        const size_t maxlong = (size_t)(nnz / (FIX_SHORT + 1)) + 1;   // rows longer than FIX_SHORT
        if ((st = lists.alloc(2 * maxlong)) != CSB200_OK) return fail(st);
        if ((st = counts.alloc(2)) != CSB200_OK) return fail(st);
        TR_CUDA(cudaMemsetAsync(counts.ptr, 0, 2 * sizeof(int), s));
        int *mid = lists.ptr, *big = lists.ptr + maxlong;
        if (has_x) k_fix_short<true><<<ceil_div(m, 256), 256, 0, s>>>(m, C->p, C->i, C->x, A->p, A->i, A->x, mid, big, counts.ptr);
        else       k_fix_short<false><<<ceil_div(m, 256), 256, 0, s>>>(m, C->p, C->i, nullptr, A->p, A->i, nullptr, mid, big, counts.ptr);
        TR_LAUNCHED();
        if (nnz > FIX_SHORT) {
            const int g_mid = (int)min((long long)148 * 16, (long long)ceil_div((long long)maxlong, 4));
            if (has_x) k_fix_mid<true><<<g_mid, 128, 0, s>>>(mid, counts.ptr, C->p, C->i, C->x, A->p, A->i, A->x);
            else       k_fix_mid<false><<<g_mid, 128, 0, s>>>(mid, counts.ptr, C->p, C->i, nullptr, A->p, A->i, nullptr);
            TR_LAUNCHED();
        }
        if (nnz > FIX_MID) {
            const int g_big = 148 * 4;
            if (has_x) k_fix_big<true><<<g_big, 512, 0, s>>>(big, counts.ptr, C->p, C->i, C->x, A->p, A->i, A->x);
            else       k_fix_big<false><<<g_big, 512, 0, s>>>(big, counts.ptr, C->p, C->i, nullptr, A->p, A->i, nullptr);
            TR_LAUNCHED();
        }
    }
#undef TR_CUDA
#undef TR_LAUNCHED
    *out = C;
    return CSB200_OK;
}

}  // namespace csb
