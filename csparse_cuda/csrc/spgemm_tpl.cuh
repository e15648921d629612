// spgemm_tpl.cuh -- cs_multiply on PATTERN CLASSES (included by spgemm.cu).
//
// Gustavson's column j of C = A*B (csparse.py:1629-1639) is a function of three things only: the
// rows k of B(:,j) RELATIVE to j, the patterns of the columns A(:,k) RELATIVE to k, and the values.
// Two columns j, j' for which the first two agree have the same pattern up to the shift j' - j,
// discover their rows in the same order and send every product to the same position of the column.
// Matrices of translation-invariant operators (stencils on grids, banded / Toeplitz-like matrices)
// have a handful of such classes -- the 27-point stencil on 128^3 has 27 classes of A columns and 125
// classes of columns of A*A for 2.1 M columns.  For those the symbolic phase (cs_scatter's mark test,
// csparse.py:1979-1988) runs ONCE per class on a representative column and leaves a template:
//     cnt                 rows of the column
//     rows[t] - j         the rows in the reference's discovery order, relative to the column
//     pos[q]              position in the column of the q-th product (B storage order, then A storage order)
// and every other column of the class is pure arithmetic: vals[pos[q]] += B(k,j) * A(i,k) with no
// hash table, no row index loaded and no per-column symbolic work.  Sums are formed in the reference's
// order and rows come out in the reference's discovery order, so p, i, x are bit-identical to
// cs_multiply.
//
// Classes are found by hashing (64 bit, position-dependent) into a small open-addressing table and
// then VERIFIED entry by entry against the class representative, so a hash collision only sends a
// column to the general kernels.  Matrices without such structure overflow the table (CLS_MAX classes)
// within the first few thousand columns; every later warp sees the abort flag and the cost is one pass
// over the row indices.
#pragma once

namespace csb {

constexpr int CLS_SLOTS = 2048;          // open-addressing table
constexpr int CLS_MAX = 1024;            // more classes than this: not a structured matrix
constexpr int TPL_UB = 4096;             // products per column a template holds
constexpr int TPL_CAP = 256;             // rows per column a template holds (positions fit a byte)
constexpr int TPL_MIN_N = 16384;         // below this the extra launches cost more than they save

struct ClsTable {
    unsigned long long *keys;            // CLS_SLOTS, 0 = empty
    int *rep;                            // CLS_SLOTS: a member column (the smallest of those that touched the table)
    int *dense;                          // CLS_SLOTS: slot -> dense class id
    int *rep_dense;                      // CLS_MAX: dense class id -> representative column
    int *info;                           // [0] classes inserted [1] aborted [2] dense classes [3] templated columns
};
constexpr size_t CLS_TABLE_BYTES = CLS_SLOTS * 8 + CLS_SLOTS * 4 + CLS_SLOTS * 4 + CLS_MAX * 4 + 64;

static inline ClsTable cls_table_at(void *base)
{
    ClsTable t;
    char *b = static_cast<char *>(base);
    t.keys = reinterpret_cast<unsigned long long *>(b);
    t.rep = reinterpret_cast<int *>(b + CLS_SLOTS * 8);
    t.dense = t.rep + CLS_SLOTS;
    t.rep_dense = t.dense + CLS_SLOTS;
    t.info = t.rep_dense + CLS_MAX;
    return t;
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ unsigned long long warp_sum64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void k_cls_init(ClsTable t)
{
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < CLS_SLOTS) { t.keys[s] = 0ull; t.rep[s] = INT_MAX; t.dense[s] = -1; }
    if (s < 16) t.info[s] = 0;
}

// one lane of the warp: slot of the class with hash h (created if new), -1 once the table gave up.
// Reads come from L2 and the atomics are issued only when they can change something: two million
// columns of one class must not queue on one address.
__device__ __forceinline__ int cls_insert(const ClsTable &t, unsigned long long h, int col)
{
    unsigned slot = (unsigned)(h >> 40) & (CLS_SLOTS - 1);
    for (int probe = 0; probe < CLS_SLOTS; probe++) {
        if (__ldcg(t.info + 1)) return -1;
        unsigned long long k = __ldcg(t.keys + slot);
        if (k == 0ull) {
            k = atomicCAS(t.keys + slot, 0ull, h);
            if (k == 0ull) {
                if (atomicAdd(t.info, 1) >= CLS_MAX) { atomicExch(t.info + 1, 1); return -1; }
                k = h;
            }
        }
        if (k == h) {
            if (col < __ldcg(t.rep + slot)) atomicMin(t.rep + slot, col);
            return (int)slot;
        }
        slot = (slot + 1) & (CLS_SLOTS - 1);
    }
    atomicExch(t.info + 1, 1);
    return -1;
}

// ---- classes of the columns of A: the rows relative to the column index, in storage order ------
// Resident warps stride over the columns and remember the class of the previous column: in a
// structured matrix nearly every column repeats it, and two million warps reading one table slot
// would queue on a single L2 sector.
constexpr int CLS_WARPS = 8;
__global__ void __launch_bounds__(CLS_WARPS * 32)
k_cls_hash_a(int n, const csi *__restrict__ Ap, const csi *__restrict__ Ai, ClsTable t, int *__restrict__ ca)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * CLS_WARPS;
    unsigned long long last_h = 0ull;
    int last_slot = -1, it = 0;
    for (int k = blockIdx.x * CLS_WARPS + (threadIdx.x >> 5); k < n; k += nwarps, it++) {
        if ((it & 15) == 0 && __ldcg(t.info + 1)) return;
        const int b = Ap[k], e = Ap[k + 1];
        unsigned long long h = 0;
        for (int p = b + lane; p < e; p += 32)
            h += mix64(mix64((unsigned long long)(p - b) + 1) ^ (unsigned long long)(unsigned)(Ai[p] - k));
        h = warp_sum64(h) + mix64(0xA5A5A5A5ull + (unsigned long long)(e - b));
        if (h == 0ull) h = 1ull;
        if (lane == 0) {
            if (h != last_h) { last_slot = cls_insert(t, h, k); last_h = last_slot >= 0 ? h : 0ull; }
            ca[k] = last_slot;
        }
    }
}

// dense class ids (slot order), one CTA of 1024 threads
__global__ void __launch_bounds__(1024) k_cls_compact(ClsTable t)
{
    __shared__ int warp_tot[32];
    static_assert(CLS_SLOTS == 2048, "two slots per thread");
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (__ldcg(t.info + 1)) { if (tid == 0) t.info[2] = 0; return; }
    const int f0 = t.keys[2 * tid] != 0ull, f1 = t.keys[2 * tid + 1] != 0ull;
    int inc = f0 + f1;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += u; }
        warp_tot[lane] = w;
    }
    __syncthreads();
    int id = inc - (f0 + f1) + (wid ? warp_tot[wid - 1] : 0);
    if (f0) { t.dense[2 * tid] = id; t.rep_dense[id] = t.rep[2 * tid]; id++; }
    if (f1) { t.dense[2 * tid + 1] = id; t.rep_dense[id] = t.rep[2 * tid + 1]; }
    if (tid == 1023) t.info[2] = warp_tot[31];
}

// entry-by-entry comparison with the representative; ca[k] <- dense class id, or -1
__global__ void __launch_bounds__(256)
k_cls_verify_a(int n, const csi *__restrict__ Ap, const csi *__restrict__ Ai, ClsTable t, int *__restrict__ ca)
{
    const int lane = threadIdx.x & 31;
    const int k = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (k >= n) return;
    if (t.info[1]) return;              // final by now: a cached load
    const int slot = ca[k];
    bool ok = slot >= 0;
    int r = 0;
    if (ok) r = t.rep[slot];
    if (ok && r != k) {
        const int b = Ap[k], len = Ap[k + 1] - b, br = Ap[r];
        ok = (Ap[r + 1] - br) == len;
        if (ok)
            for (int e = lane; e < len; e += 32)
                if (Ai[b + e] - k != Ai[br + e] - r) ok = false;
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    if (lane == 0) ca[k] = ok ? t.dense[slot] : -1;
}

// ---- classes of the columns of B: (row relative to the column, class of that column of A) ------
__global__ void __launch_bounds__(CLS_WARPS * 32)
k_cls_hash_b(int n, const csi *__restrict__ Bp, const csi *__restrict__ Bi, const int *__restrict__ ca,
             ClsTable t, int *__restrict__ cb)
{
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * CLS_WARPS;
    unsigned long long last_h = 0ull;
    int last_slot = -1, it = 0;
    for (int j = blockIdx.x * CLS_WARPS + (threadIdx.x >> 5); j < n; j += nwarps, it++) {
        if ((it & 15) == 0 && __ldcg(t.info + 1)) return;
        const int b = Bp[j], e = Bp[j + 1];
        unsigned long long h = 0;
        bool ok = e > b;
        for (int p = b + lane; p < e; p += 32) {
            const int k = Bi[p];
            const int c = ca[k];
            if (c < 0) ok = false;
            h += mix64(mix64((unsigned long long)(p - b) + 1) ^ (unsigned long long)(unsigned)(k - j) ^
                       ((unsigned long long)(unsigned)c << 32));
        }
        ok = __all_sync(0xffffffffu, ok);
        h = warp_sum64(h) + mix64(0x5A5A5A5Aull + (unsigned long long)(e - b));
        if (h == 0ull) h = 1ull;
        if (lane == 0) {
            int slot = -1;
            if (ok) {
                if (h != last_h) { last_slot = cls_insert(t, h, j); last_h = last_slot >= 0 ? h : 0ull; }
                slot = last_slot;
            }
            cb[j] = slot;
        }
    }
}

__global__ void __launch_bounds__(256)
k_cls_verify_b(int n, const csi *__restrict__ Bp, const csi *__restrict__ Bi, const int *__restrict__ ca,
               ClsTable t, int *__restrict__ cb)
{
    const int lane = threadIdx.x & 31;
    const int j = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (j >= n) return;
    if (t.info[1]) return;              // final by now: a cached load
    const int slot = cb[j];
    bool ok = slot >= 0;
    int r = 0;
    if (ok) r = t.rep[slot];
    if (ok && r != j) {
        const int b = Bp[j], len = Bp[j + 1] - b, br = Bp[r];
        ok = (Bp[r + 1] - br) == len;
        if (ok)
            for (int e = lane; e < len; e += 32) {
                const int k = Bi[b + e], kr = Bi[br + e];
                if (k - j != kr - r || ca[k] != ca[kr]) ok = false;
            }
    }
    ok = __all_sync(0xffffffffu, ok);
    __syncwarp();
    if (lane == 0) cb[j] = ok ? t.dense[slot] : -1;
}

// ---- the template of a class: cs_scatter on its representative column, one warp ------------------
// A's columns are canonical (distinct rows per column), so the rows of a 32-entry step are distinct.
__global__ void __launch_bounds__(32)
k_tpl_build(ClsTable t, const csi *__restrict__ Ap, const csi *__restrict__ Ai, const csi *__restrict__ Bp,
            const csi *__restrict__ Bi, const int *__restrict__ ub, int *__restrict__ tpl_cnt,
            unsigned char *__restrict__ tpl_pos, int *__restrict__ tpl_rows)
{
    constexpr int LOGH = 10, H = 1 << LOGH;
    static_assert(H >= 2 * (TPL_CAP + 32), "the table never fills");
    __shared__ int keys[H];
    __shared__ unsigned short posof[H];
    const int lane = threadIdx.x, c = blockIdx.x;
    if (__ldcg(t.info + 1) || c >= __ldcg(t.info + 2)) return;
    const int j = t.rep_dense[c];
    if (ub[j] > TPL_UB) { if (lane == 0) tpl_cnt[c] = -1; return; }
    for (int s = lane; s < H; s += 32) keys[s] = EMPTY;
    __syncwarp();
    const unsigned lt = lanemask_lt();
    int cnt = 0, q0 = 0;
    bool fail = false;
    const int pb_end = Bp[j + 1];
    for (int pb = Bp[j]; pb < pb_end && !fail; pb++) {
        const int k = Bi[pb];
        const int ab = Ap[k], ae = Ap[k + 1];
        for (int pa0 = ab; pa0 < ae; pa0 += 32) {
            const int pa = pa0 + lane;
            const bool active = pa < ae;
            const int i = active ? Ai[pa] : 0;
            bool isnew;
            const int slot = warp_find_or_insert(keys, H - 1, i, active, hash_row(i, LOGH), isnew);
            const unsigned newmask = __ballot_sync(0xffffffffu, isnew);
            if (isnew) {
                const int pos = cnt + __popc(newmask & lt);
                posof[slot] = (unsigned short)pos;
                if (pos < TPL_CAP) tpl_rows[(size_t)c * TPL_CAP + pos] = i - j;
            }
            __syncwarp();
            if (active) tpl_pos[(size_t)c * TPL_UB + q0 + (pa - ab)] = (unsigned char)posof[slot];
            cnt += __popc(newmask);
            if (cnt > TPL_CAP) { fail = true; break; }
        }
        q0 += ae - ab;
    }
    if (lane == 0) tpl_cnt[c] = fail ? -1 : cnt;
}

// columns of a class with a template: cnt[j] is known, the general symbolic phase skips them (ub 0)
__global__ void k_tpl_apply(int n, ClsTable t, const int *__restrict__ tpl_cnt, int *__restrict__ cb,
                            int *__restrict__ cnt, int *__restrict__ ub)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = !t.info[1];
    bool hit = false;
    if (j < n) {
        const int c = live ? cb[j] : -1;
        const int tc = c >= 0 ? tpl_cnt[c] : -1;
        hit = tc >= 0;
        if (hit) { cnt[j] = tc; ub[j] = 0; } else cb[j] = -1;
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(t.info + 3, __popc(m));
}

// ---- numeric on templates, one warp per column ----------------------------------------------------
// Shared memory per warp: TPL_CAP accumulators + the staged (first, length, B value) of 32 columns of A.
// The kernel is latency-bound unless loads are kept in flight: the values of TPL_BATCH columns of A
// (and their positions) are fetched together before the read-modify-write steps that consume them,
// and the B entries of the warp's NEXT column are fetched while the current one is accumulated.
constexpr int TPL_PER_WARP = TPL_CAP * 8 + 32 * 16;
template <bool VALUES, int TPL_BATCH, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_num_tpl(int n, const int *__restrict__ cb, const int *__restrict__ tpl_cnt,
          const unsigned char *__restrict__ tpl_pos, const int *__restrict__ tpl_rows,
          const csi *__restrict__ Ap, const double *__restrict__ Ax,
          const csi *__restrict__ Bp, const csi *__restrict__ Bi, const double *__restrict__ Bx,
          const csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx)
{
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    double *vals = reinterpret_cast<double *>(sm_raw + (size_t)wid * TPL_PER_WARP);
    int4 *stage = reinterpret_cast<int4 *>(sm_raw + (size_t)wid * TPL_PER_WARP + TPL_CAP * 8);
    const int nwarps = gridDim.x * 8;

    // (first entry, length, B value) of the A column named by entry pb of B, as one int4
    auto fetch = [&](int pb, int pb_end) -> int4 {
        int4 st = make_int4(0, 0, 0, 0);
        if (pb < pb_end) {
            const int k = Bi[pb];
            const int ab = Ap[k], ae = Ap[k + 1];
            const double beta = VALUES ? Bx[pb] : 0.0;
            st = make_int4(ab, ae - ab, __double2loint(beta), __double2hiint(beta));
        }
        return st;
    };

    int j = blockIdx.x * 8 + wid;
    int c = -1, pb_begin = 0, pb_end = 0;
    int4 st = make_int4(0, 0, 0, 0);
    if (j < n) {
        c = cb[j];
        if (c >= 0) { pb_begin = Bp[j]; pb_end = Bp[j + 1]; if (VALUES) st = fetch(pb_begin + lane, pb_end); }
    }
    while (j < n) {
        // the next column of this warp: its class and first 32 B entries travel while this one is summed
        const int jn = j + nwarps;
        int cn = -1, pbn_begin = 0, pbn_end = 0;
        int4 stn = make_int4(0, 0, 0, 0);
        if (c >= 0) {
            const int cnt = tpl_cnt[c];
            const int out = Cp[j];
            if (VALUES) {
                // -0.0 is the exact additive identity: the first product of a row lands as the
                // reference's first-touch assignment (csparse.py:1986)
                for (int t = lane; t < cnt; t += 32) vals[t] = -0.0;
                const unsigned char *pm = tpl_pos + (size_t)c * TPL_UB;
                for (int pb0 = pb_begin; pb0 < pb_end; pb0 += 32) {
                    if (pb0 != pb_begin) st = fetch(pb0 + lane, pb_end);
                    __syncwarp();
                    stage[lane] = st;
                    __syncwarp();
                    if (pb0 == pb_begin && jn < n) {
                        cn = cb[jn];
                        if (cn >= 0) { pbn_begin = Bp[jn]; pbn_end = Bp[jn + 1]; stn = fetch(pbn_begin + lane, pbn_end); }
                    }
                    const int nb = min(32, pb_end - pb0);
                    for (int s0 = 0; s0 < nb; s0 += TPL_BATCH) {
                        double prod[TPL_BATCH];
                        int pos[TPL_BATCH];
                        unsigned act = 0;
                        bool longcol = false;
                        const unsigned char *pm0 = pm;
#pragma unroll
                        for (int u = 0; u < TPL_BATCH; u++) {
                            const int4 g = s0 + u < nb ? stage[s0 + u] : make_int4(0, 0, 0, 0);   // broadcast LDS.128
                            prod[u] = 0.0;
                            pos[u] = 0;
                            if (lane < g.y) {
                                prod[u] = Ax[g.x + lane];                       // TPL_BATCH loads in flight
                                pos[u] = pm[lane];
                                act |= 1u << u;
                            }
                            pm += g.y;
                            longcol |= g.y > 32;
                        }
                        if (!longcol) {
#pragma unroll
                            for (int u = 0; u < TPL_BATCH; u++) {
                                if (act & (1u << u)) {
                                    const int2 bw = *reinterpret_cast<const int2 *>(&stage[s0 + u].z);   // B(k,j) again: registers are scarce
                                    vals[pos[u]] = __dadd_rn(vals[pos[u]], __dmul_rn(__hiloint2double(bw.y, bw.x), prod[u]));   // csparse.py:1988
                                }
                                __syncwarp();
                            }
                        } else {
                            // a column of A longer than a warp: the same steps, 32 entries at a time
                            const unsigned char *pq = pm0;
                            for (int u = 0; u < TPL_BATCH && s0 + u < nb; u++) {
                                const int4 g = stage[s0 + u];
                                const double beta = __hiloint2double(g.w, g.z);
                                for (int o0 = 0; o0 < g.y; o0 += 32) {
                                    const int e = o0 + lane;
                                    if (e < g.y) {
                                        const int ps = pq[e];
                                        vals[ps] = __dadd_rn(vals[ps], __dmul_rn(beta, Ax[g.x + e]));
                                    }
                                    __syncwarp();
                                }
                                pq += g.y;
                            }
                        }
                    }
                }
            } else if (jn < n) {
                cn = cb[jn];
            }
            const int *rows = tpl_rows + (size_t)c * TPL_CAP;
            for (int t = lane; t < cnt; t += 32) {
                Ci[out + t] = j + rows[t];
                if (VALUES) Cx[out + t] = vals[t];
            }
            __syncwarp();
        } else if (jn < n) {
            cn = cb[jn];
            if (VALUES && cn >= 0) { pbn_begin = Bp[jn]; pbn_end = Bp[jn + 1]; stn = fetch(pbn_begin + lane, pbn_end); }
        }
        j = jn; c = cn; pb_begin = pbn_begin; pb_end = pbn_end; st = stn;
    }
}

}  // namespace csb
