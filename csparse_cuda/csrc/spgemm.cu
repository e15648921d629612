// spgemm.cu -- cs_multiply (csparse.py:1608-1642) and its inner kernel cs_scatter
// (csparse.py:1961-1989): C = A*B, column by column (Gustavson), as a two-phase
// SpGEMM:
//
//   k_ub        per column j of B: ub[j] = sum_k nnz(A(:,k)) over k in B(:,j)
//               (the number of multiply-adds; also the size bound used for binning)
//   k_bin       columns -> work lists by size class (no host round trip per class)
//   k_sym_*     symbolic: cnt[j] = |union of the patterns A(:,k)|  -- structural,
//               explicit and cancelled zeros count, exactly like cs_scatter's mark test
//   excl_scan   Cp = cumsum(cnt) (scan.cu); replaces the reference's realloc doubling
//   k_num_*     numeric: hash-accumulate x[i] += B(k,j)*A(i,k), emit (Ci, Cx)
//
// The reference's dense workspaces w[m] (marks) and x[m] (accumulator) become a
// per-warp hash table in shared memory (row -> discovery position) plus dense
// per-position arrays; columns too large for shared memory use the reference's
// own dense w[m]/x[m] scheme in global memory, one workspace per resident CTA.
//
// Order and rounding: a warp walks B(:,j) in storage order and each A(:,k) in
// storage order (32 entries per step, lane = storage offset), products are
// rounded before they are added, first touch assigns.  New rows are ranked by
// ballot/popc, so C's columns come out in the reference's DISCOVERY ORDER and the
// sums are formed in the reference's order: p, i and x are bit-identical to
// cs_multiply, not merely equal after a sort.  (A with duplicate entries inside a
// column makes two lanes hit one row in a step; that case is serialised in lane
// order, see CANON.)
#include "common.cuh"
#include <type_traits>
#include <stdlib.h>

namespace csb {

constexpr int EMPTY = -1;
constexpr int BLK_STRIDE = 64;          // (block, mask) pairs kept per column for the blocked numeric kernel
constexpr int BLK_CAP = 128;            // rows per column the blocked numeric kernel holds
constexpr int BLK_LOGH = 7, BLK_H = 1 << BLK_LOGH;   // slots of its block table (2 * BLK_STRIDE); 3 KB per warp: 8 CTAs per SM

__device__ __forceinline__ unsigned hash_row(int i, int logh)
{
    return ((unsigned)i * 0x9E3779B1u) >> (32 - logh);
}

// ---- sizes and bins ------------------------------------------------------------
__global__ void k_ub(int nB, const csi *__restrict__ Bp, const csi *__restrict__ Bi,
                     const csi *__restrict__ Ap, int *__restrict__ ub, unsigned long long *flops)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    long long s = 0;
    if (j < nB) {
        for (int p = Bp[j]; p < Bp[j + 1]; p++) {
            const int k = Bi[p];
            s += Ap[k + 1] - Ap[k];
        }
        ub[j] = (int)min(s, (long long)INT_MAX);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(flops, (unsigned long long)s);
}

// size[j] -> class by thresholds t0 < t1 < t2: 0 -> skipped, (0,t0] -> list 0,
// (t0,t1] -> list 1, (t1,t2] -> list 2, > t2 -> list 3
__global__ void k_bin(int n, const int *__restrict__ size, int cap, int t0, int t1, int t2,
                      int *__restrict__ lists, int *__restrict__ counts)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    int b = -1;
    if (j < n) {
        const int s = min(size[j], cap);
        if (s > 0) b = s <= t0 ? 0 : s <= t1 ? 1 : s <= t2 ? 2 : 3;
    }
    // one atomic per warp and class instead of one per column
    const unsigned lt = lanemask_lt();
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const unsigned mask = __ballot_sync(0xffffffffu, b == c);
        if (mask) {
            int base = 0;
            const int leader = __ffs(mask) - 1;
            if ((threadIdx.x & 31) == leader) base = atomicAdd(&counts[c], __popc(mask));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (b == c) lists[(size_t)c * n + base + __popc(mask & lt)] = j;
        }
    }
}

// ---- warp-private open-addressing insert/lookup without atomics ---------------------
// One warp owns the table, so a racing plain store + read-back replaces atomicCAS
// (ATOMS costs ~2 cycles per lane, a conflict-free LDS ~1 cycle per warp): every
// lane that still looks for row i reads keys[h]; a hit ends its search; lanes that
// saw EMPTY all store their row, the warp synchronises, and whoever reads its own
// row back owns the slot -- the losers (distinct rows, same slot) move on.  Rows of
// the lanes with want == true must be pairwise distinct.  Returns the slot.
__device__ __forceinline__ int warp_find_or_insert(int *keys, int hmask, int i, bool want,
                                                   unsigned h, bool &isnew)
{
    bool pend = want;
    isnew = false;
    while (true) {
        int k = EMPTY - 1;
        if (pend) k = keys[h];
        if (k == i) pend = false;
        const bool tryins = pend && k == EMPTY;
        if (__any_sync(0xffffffffu, tryins)) {
            if (tryins) keys[h] = i;
            __syncwarp();
            if (tryins && keys[h] == i) { isnew = true; pend = false; }
        }
        if (!__any_sync(0xffffffffu, pend)) break;
        if (pend) h = (h + 1) & hmask;
    }
    return (int)h;
}

}  // namespace csb
#include "spgemm_tpl.cuh"
namespace csb {

// ---- compressed columns for the symbolic phase ------------------------------------------
// Column k of A as pairs (block = row >> 5, mask = bits of the rows present in that block),
// runs of equal blocks merged (sorted columns of a stencil shrink ~3x).  The pairs of column k
// live at offset Ap[k] of blk / mask (no second pointer array), len32[k] of them.
__global__ void __launch_bounds__(256)
k_compress_cols(int n, const csi *__restrict__ Ap, const csi *__restrict__ Ai,
                csi *__restrict__ blk32, unsigned *__restrict__ mask32, csi *__restrict__ len32)
{
    const int lane = threadIdx.x & 31;
    const int k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n) return;
    const int b = Ap[k], e = Ap[k + 1];
    int out = 0;                                               // warp-uniform: pairs written so far
    for (int p0 = b; p0 < e; p0 += 32) {
        const int p = p0 + lane;
        const bool valid = p < e;
        const int row = valid ? Ai[p] : 0;
        const int blk = valid ? (row >> 5) : -1 - lane;        // past the end: matches no neighbour
        unsigned v = valid ? 1u << (row & 31) : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {                     // segmented inclusive OR over runs
            const unsigned u = __shfl_up_sync(0xffffffffu, v, o);
            const int bu = __shfl_up_sync(0xffffffffu, blk, o);
            if (lane >= o && bu == blk) v |= u;
        }
        const int bnext = __shfl_down_sync(0xffffffffu, blk, 1);
        const bool tail = valid && (lane == 31 || bnext != blk);
        const unsigned tails = __ballot_sync(0xffffffffu, tail);
        if (tail) {
            const int q = b + out + __popc(tails & lanemask_lt());
            blk32[q] = blk;
            mask32[q] = v;
        }
        out += __popc(tails);
    }
    if (lane == 0) len32[k] = out;
}

// ---- symbolic on compressed columns, one warp per column of B ------------------------------
// The (block, mask) pairs of all A(:,k), k in B(:,j), are taken 32 at a time regardless of
// which A column they come from (flattened), so the lanes stay busy for short columns.  The
// hash set is keyed by block; masks are OR-ed in and the newly set bits counted.  Two lanes may
// carry the same block in one step, hence shared-memory atomics (CAS insert, OR); the count is
// a set size, so the order does not matter.  Tables are emptied by undoing the touched slots.
template <int LOGH, int CAP, int WARPS, bool OPTIMISTIC>
__global__ void __launch_bounds__(WARPS * 32)
k_sym_flat(const int *__restrict__ list, const int *__restrict__ ncols_dev, int ncols_host,
           const csi *__restrict__ Ap, const csi *__restrict__ blk32, const unsigned *__restrict__ mask32,
           const csi *__restrict__ len32,
           const csi *__restrict__ Bp, const csi *__restrict__ Bi, int *__restrict__ cnt_out,
           int *__restrict__ ovf_list, int *ovf_count, int2 *__restrict__ blk_out, int *__restrict__ nblk_out)
{
    constexpr int H = 1 << LOGH;
    static_assert(H >= CAP + 64 && H <= 65536, "table must never fill");
    constexpr int PER_WARP = H * 8 + CAP * 2;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int *keys = reinterpret_cast<int *>(sm_raw + (size_t)wid * PER_WARP);
    unsigned *masks = reinterpret_cast<unsigned *>(keys + H);
    unsigned short *slots = reinterpret_cast<unsigned short *>(masks + H);
    const unsigned lt = lanemask_lt();
    const int ncols = ncols_dev ? *ncols_dev : ncols_host;
    const int nwarps = gridDim.x * WARPS;
    for (int s = lane; s < H; s += 32) { keys[s] = EMPTY; masks[s] = 0; }
    __syncwarp();
    for (int idx = blockIdx.x * WARPS + wid; idx < ncols; idx += nwarps) {
        const int j = list[idx];
        int nslots = 0;                                // warp-uniform: distinct blocks so far
        int mine = 0;                                  // rows first seen by this lane
        bool ovf = false;
        const int pb_end = Bp[j + 1];
        for (int pb0 = Bp[j]; pb0 < pb_end && !ovf; pb0 += 32) {
            int base = 0, len = 0;
            if (pb0 + lane < pb_end) {
                const int k = Bi[pb0 + lane];
                base = Ap[k];
                len = len32[k];
            }
            int incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - len;
            for (int f0 = 0; f0 < total; f0 += 32) {
                const int f = f0 + lane;
                const bool valid = f < total;
                int s = 0;                             // source lane: the largest s with excl[s] <= f
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int es = __shfl_sync(0xffffffffu, excl, (s + step) & 31);
                    if (es <= f) s += step;            // s + step <= 31 always
                }
                const int sb = __shfl_sync(0xffffffffu, base, s);
                const int se = __shfl_sync(0xffffffffu, excl, s);
                int slot = 0;
                bool isnew = false;
                if (valid) {
                    const int blk = blk32[sb + (f - se)];
                    const unsigned m = mask32[sb + (f - se)];
                    unsigned h = hash_row(blk, LOGH);
                    while (true) {
                        const int old = atomicCAS(&keys[h], EMPTY, blk);
                        if (old == EMPTY) { isnew = true; break; }
                        if (old == blk) break;
                        h = (h + 1) & (H - 1);
                    }
                    slot = (int)h;
                    const unsigned before = atomicOr(&masks[slot], m);
                    mine += __popc(m & ~before);
                }
                const unsigned newmask = __ballot_sync(0xffffffffu, isnew);
                if (newmask) {
                    const int pos = nslots + __popc(newmask & lt);
                    __syncwarp();                                           // this step's ORs have landed
                    if (isnew) {
                        if (pos < CAP) slots[pos] = (unsigned short)slot;
                        else { keys[slot] = EMPTY; masks[slot] = 0; }      // beyond the undo list: undo now
                    }
                    nslots += __popc(newmask);
                    if (nslots > CAP) { ovf = true; break; }
                }
            }
        }
        __syncwarp();
        const int filled = min(nslots, CAP);
        if (blk_out) {
            // the column's (block, mask) set in discovery order, for the blocked numeric kernel
            const bool keep = !ovf && nslots <= BLK_STRIDE;
            if (keep)
                for (int t = lane; t < nslots; t += 32) {
                    const int sl = slots[t];
                    blk_out[(size_t)j * BLK_STRIDE + t] = make_int2(keys[sl], (int)masks[sl]);
                }
            if (lane == 0 && !ovf) nblk_out[j] = keep ? nslots : -1;
        }
        for (int t = lane; t < filled; t += 32) { const int sl = slots[t]; keys[sl] = EMPTY; masks[sl] = 0; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if (lane == 0) {
            if (!ovf) cnt_out[j] = mine;
            else if (OPTIMISTIC) ovf_list[atomicAdd(ovf_count, 1)] = j;
        }
        __syncwarp();
    }
}

// ---- symbolic, one CTA per column, dense marks in global memory --------------------
// marks[ws][m] plays the reference's w[] (csparse.py:1624): zero-initialised once,
// never cleared; the mark of a column is a per-workspace counter + 1.
__global__ void __launch_bounds__(256)
k_sym_dense(const int *__restrict__ list, int ncols, int m,
            const csi *__restrict__ Ap, const csi *__restrict__ Ai,
            const csi *__restrict__ Bp, const csi *__restrict__ Bi,
            int *marks, int *__restrict__ cnt_out)
{
    __shared__ int s_cnt;
    int *w = marks + (size_t)blockIdx.x * m;
    int mark = 0;
    for (int idx = blockIdx.x; idx < ncols; idx += gridDim.x) {
        const int j = list[idx];
        mark++;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int mine = 0;
        for (int pb = Bp[j]; pb < Bp[j + 1]; pb++) {
            const int k = Bi[pb];
            for (int pa = Ap[k] + threadIdx.x; pa < Ap[k + 1]; pa += blockDim.x)
                if (atomicMax(&w[Ai[pa]], mark) < mark) mine++;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
        if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_cnt, mine);
        __syncthreads();
        if (threadIdx.x == 0) cnt_out[j] = s_cnt;
        __syncthreads();
    }
}

// ---- numeric, one warp per column ---------------------------------------------------
// Shared memory per warp: vals[H] (accumulator of the row held by that slot), keys[H]
// (row or EMPTY), order[CAP] (slot of the t-th discovered row).  The table is emptied
// while the column is emitted.  CANON: every column of A has distinct rows.
template <int LOGH, int CAP, int WARPS, bool VALUES, bool CANON>
__global__ void __launch_bounds__(WARPS * 32)
k_num_warp(const int *__restrict__ list, int ncols,
           const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
           const csi *__restrict__ Bp, const csi *__restrict__ Bi, const double *__restrict__ Bx,
           const csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx)
{
    constexpr int H = 1 << LOGH;
    static_assert(H >= 2 * CAP && H <= 65536, "load factor <= 1/2, slots fit 16 bits");
    constexpr int PER_WARP = H * 8 + H * 4 + CAP * 2;      // bytes
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *base = sm_raw + (size_t)wid * PER_WARP;
    double *vals = reinterpret_cast<double *>(base);                       // doubles first (alignment)
    int *keys = reinterpret_cast<int *>(base + H * 8);
    unsigned short *order = reinterpret_cast<unsigned short *>(keys + H);
    const unsigned lt = lanemask_lt();
    const int nwarps = gridDim.x * WARPS;

    for (int s = lane; s < H; s += 32) keys[s] = EMPTY;
    __syncwarp();
    for (int idx = blockIdx.x * WARPS + wid; idx < ncols; idx += nwarps) {
        const int j = list[idx];
        int cnt = 0;                                   // warp-uniform: rows discovered so far
        const int pb_end = Bp[j + 1];
        for (int pb0 = Bp[j]; pb0 < pb_end; pb0 += 32) {
            const int my_pb = pb0 + lane;
            int my_ab = 0, my_ae = 0;
            double my_beta = 1.0;                      // pattern-only B: coefficient 1 (csparse.py:1636)
            if (my_pb < pb_end) {
                const int k = Bi[my_pb];
                my_ab = Ap[k];
                my_ae = Ap[k + 1];
                if (VALUES) my_beta = Bx[my_pb];
            }
            const int nb = min(32, pb_end - pb0);
            for (int s = 0; s < nb; s++) {
                const int ab = __shfl_sync(0xffffffffu, my_ab, s);
                const int ae = __shfl_sync(0xffffffffu, my_ae, s);
                const double beta = __shfl_sync(0xffffffffu, my_beta, s);
                for (int pa0 = ab; pa0 < ae; pa0 += 32) {          // warp-uniform trip count
                    const int pa = pa0 + lane;
                    const bool active = pa < ae;
                    int i = 0;
                    double prod = 0.0;
                    if (active) {
                        i = Ai[pa];
                        if (VALUES) prod = __dmul_rn(beta, Ax[pa]);
                    }
                    bool leader = active;              // lane that performs the insert for its row
                    int lead_lane = lane;
                    if (!CANON) {
                        const unsigned act = __ballot_sync(0xffffffffu, active);
                        if (active) {
                            lead_lane = __ffs(__match_any_sync(act, i)) - 1;
                            leader = lead_lane == lane;
                        }
                    }
                    bool isnew;
                    int slot = warp_find_or_insert(keys, H - 1, i, leader, hash_row(i, LOGH), isnew);
                    const unsigned newmask = __ballot_sync(0xffffffffu, isnew);
                    if (isnew) {
                        order[cnt + __popc(newmask & lt)] = (unsigned short)slot;
                        if (VALUES) vals[slot] = prod;             // first touch assigns (csparse.py:1986)
                    }
                    cnt += __popc(newmask);
                    if (VALUES) {
                        if (CANON) {
                            if (active && !isnew) vals[slot] = __dadd_rn(vals[slot], prod);   // (csparse.py:1988)
                        } else {
                            __syncwarp();
                            slot = __shfl_sync(0xffffffffu, slot, lead_lane);
                            unsigned pend = __ballot_sync(0xffffffffu, active && !isnew);
                            while (pend) {                          // storage (= lane) order
                                const int l = __ffs(pend) - 1;
                                if (lane == l) vals[slot] = __dadd_rn(vals[slot], prod);
                                pend &= pend - 1;
                                __syncwarp();
                            }
                        }
                    }
                    __syncwarp();
                }
            }
        }
        const int out = Cp[j];
        for (int t = lane; t < cnt; t += 32) {
            const int slot = order[t];
            Ci[out + t] = keys[slot];
            if (VALUES) Cx[out + t] = vals[slot];
            keys[slot] = EMPTY;
        }
        __syncwarp();
    }
}

// ---- numeric on the symbolic phase's block tables, one warp per column ------------------------
// The symbolic phase leaves the column's pattern as (32-row block, mask) pairs.  With that known,
// the numeric phase needs no inserts, no first-touch bookkeeping and no ranks: the pairs go back
// into a small hash set keyed by block, every block gets the offset of its first row (prefix sum
// of the mask popcounts), and a product for row i lands at base[block] + popc(mask & bits below
// i).  Per multiply-add that is one probe, one popcount and one accumulate -- about 40 warp
// instructions per 32 products against 100 for k_num_warp.  Rows of a column come out block by
// block in the blocks' discovery order, ascending inside a block (not the reference's discovery
// order; the pattern is the same set, and every value is summed in the reference's sequence).
// Shared memory per warp: vals[BLK_CAP] doubles, rows[BLK_CAP] ints, slot table {key, mask, base}.
template <bool VALUES, bool CANON>
__global__ void __launch_bounds__(256)
k_num_blocked(const int *__restrict__ list, int ncols,
              const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
              const csi *__restrict__ Bp, const csi *__restrict__ Bi, const double *__restrict__ Bx,
              const int2 *__restrict__ blk_in, const int *__restrict__ nblk_in,
              const csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx)
{
    constexpr int H = BLK_H, LOGH = BLK_LOGH;
    constexpr int PER_WARP = BLK_CAP * 8 + BLK_CAP * 4 + H * 12;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *base_ptr = sm_raw + (size_t)wid * PER_WARP;
    double *vals = reinterpret_cast<double *>(base_ptr);
    int *rows = reinterpret_cast<int *>(base_ptr + BLK_CAP * 8);
    int *keys = rows + BLK_CAP;                            // H
    int2 *mb = reinterpret_cast<int2 *>(keys + H);          // H x {mask, base}
    const int nwarps = gridDim.x * 8;
    for (int s = lane; s < H; s += 32) keys[s] = EMPTY;
    __syncwarp();
    for (int idx = blockIdx.x * 8 + wid; idx < ncols; idx += nwarps) {
        const int j = list[idx];
        const int nblk = nblk_in[j];
        const int out = Cp[j];
        const int cnt = Cp[j + 1] - out;
        // ---- the block table: insert (distinct keys), offsets, row list -------------------------
        int my_slot[BLK_STRIDE / 32];
        int running = 0;
#pragma unroll
        for (int r = 0; r < BLK_STRIDE / 32; r++) {
            const int t = r * 32 + lane;
            const bool valid = t < nblk;
            int2 bm = make_int2(0, 0);
            if (valid) bm = blk_in[(size_t)j * BLK_STRIDE + t];
            bool isnew;
            const int slot = warp_find_or_insert(keys, H - 1, bm.x, valid, hash_row(bm.x, LOGH), isnew);
            my_slot[r] = valid ? slot : -1;
            const int c = valid ? __popc((unsigned)bm.y) : 0;
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
            const int first = running + inc - c;
            if (valid) {
                mb[slot] = make_int2(bm.y, first);
                unsigned mk = (unsigned)bm.y;
                int q = first;
                while (mk) { const int bit = __ffs(mk) - 1; rows[q++] = (bm.x << 5) + bit; mk &= mk - 1; }
            }
            running += __shfl_sync(0xffffffffu, inc, 31);
        }
        // -0.0 is the exact additive identity (-0.0 + x == x bit for bit, also for x == -0.0), so the
        // first product of a row lands as the reference's first-touch assignment
        if (VALUES) for (int t = lane; t < cnt; t += 32) vals[t] = -0.0;
        __syncwarp();
        // ---- accumulate --------------------------------------------------------------------------
        if (VALUES) {
            const int pb_end = Bp[j + 1];
            for (int pb0 = Bp[j]; pb0 < pb_end; pb0 += 32) {
                const int my_pb = pb0 + lane;
                int my_ab = 0, my_ae = 0;
                double my_beta = 0.0;
                if (my_pb < pb_end) {
                    const int k = Bi[my_pb];
                    my_ab = Ap[k];
                    my_ae = Ap[k + 1];
                    my_beta = Bx[my_pb];
                }
                const int nb = min(32, pb_end - pb0);
                for (int s = 0; s < nb; s++) {
                    const int ab = __shfl_sync(0xffffffffu, my_ab, s);
                    const int ae = __shfl_sync(0xffffffffu, my_ae, s);
                    const double beta = __shfl_sync(0xffffffffu, my_beta, s);
                    for (int pa0 = ab; pa0 < ae; pa0 += 32) {
                        const int pa = pa0 + lane;
                        const bool active = pa < ae;
                        int pos = 0;
                        double prod = 0.0;
                        if (active) {
                            const int i = Ai[pa];
                            prod = __dmul_rn(beta, Ax[pa]);
                            const int blk = i >> 5;
                            unsigned h = hash_row(blk, LOGH);
                            while (keys[h] != blk) h = (h + 1) & (H - 1);       // present by construction
                            const int2 e = mb[h];
                            pos = e.y + __popc((unsigned)e.x & ((1u << (i & 31)) - 1u));
                        }
                        if (CANON) {
                            if (active) vals[pos] = __dadd_rn(vals[pos], prod);  // distinct rows in a step
                        } else {
                            unsigned pend = __ballot_sync(0xffffffffu, active);  // duplicates: storage (= lane) order
                            while (pend) {
                                const int l = __ffs(pend) - 1;
                                if (lane == l) vals[pos] = __dadd_rn(vals[pos], prod);
                                pend &= pend - 1;
                                __syncwarp();
                            }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        // ---- emit, empty the table -----------------------------------------------------------------
        for (int t = lane; t < cnt; t += 32) {
            Ci[out + t] = rows[t];
            if (VALUES) Cx[out + t] = vals[t];
        }
#pragma unroll
        for (int r = 0; r < BLK_STRIDE / 32; r++) if (my_slot[r] >= 0) keys[my_slot[r]] = EMPTY;
        __syncwarp();
    }
}

// ---- the blocked numeric kernel, second version ------------------------------------------------
// Same algorithm and the same results as k_num_blocked with a leaner accumulate loop (62 -> 42 warp
// instructions and ~25 % fewer shared-memory wavefronts per A column; the kernel is bound by the
// L1/shared data stage and instruction issue, not by HBM):
//   * one 8-byte table slot {block << 7 | offset of the block's first row, mask}: one LDS.64 per
//     probe instead of a key load plus a {mask, base} load (needs rows < 2^29);
//   * the (first position, length, B value) of the <= 32 A columns of a chunk of B(:,j) are staged
//     in shared memory and broadcast with one LDS.128 per A column instead of four shuffles (a
//     shuffle costs a shared-memory wavefront too);
//   * columns of A longer than a warp step take a separate copy of the loop.
constexpr int BLK2_BASE_BITS = 7;
static_assert((1 << BLK2_BASE_BITS) >= BLK_CAP, "offset field holds every row position");
constexpr long long BLK2_MAX_ROWS = 1LL << (31 - BLK2_BASE_BITS + 5);
constexpr int BLK2_PER_WARP = BLK_CAP * 8 + BLK_CAP * 4 + BLK_H * 8 + 32 * 16;

template <bool VALUES, bool CANON>
__global__ void __launch_bounds__(256, 8)
k_num_blocked2(const int *__restrict__ list, int ncols,
               const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
               const csi *__restrict__ Bp, const csi *__restrict__ Bi, const double *__restrict__ Bx,
               const int2 *__restrict__ blk_in, const int *__restrict__ nblk_in,
               const csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx)
{
    constexpr int H = BLK_H, LOGH = BLK_LOGH;
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned char *wbase = sm_raw + (size_t)wid * BLK2_PER_WARP;
    double *vals = reinterpret_cast<double *>(wbase);                               // BLK_CAP accumulators
    int *rows = reinterpret_cast<int *>(wbase + BLK_CAP * 8);                       // BLK_CAP row indices
    int2 *tab = reinterpret_cast<int2 *>(wbase + BLK_CAP * 12);                     // H x {block << 7 | base, mask}
    int4 *stage = reinterpret_cast<int4 *>(wbase + BLK_CAP * 12 + H * 8);           // 32 x {first, len, B value}
    const int nwarps = gridDim.x * 8;
    for (int s = lane; s < H; s += 32) tab[s] = make_int2(EMPTY, 0);
    __syncwarp();

    // position of row i in the column: the block's slot (present by construction), then the rank of
    // the row's bit inside the block's mask
    auto lookup = [&](int i) -> int {
        const int blk = i >> 5;
        unsigned h = hash_row(blk, LOGH);
        int2 e = tab[h];
        while ((e.x >> BLK2_BASE_BITS) != blk) { h = (h + 1) & (H - 1); e = tab[h]; }
        return (e.x & ((1 << BLK2_BASE_BITS) - 1)) + __popc((unsigned)e.y & ((1u << (i & 31)) - 1u));
    };

    // one product beta * a for row i, accumulated in the reference's order
    auto accumulate = [&](bool active, int i, double a, double beta) {
        if (CANON) {
            if (active) {                                               // distinct rows in a step
                const int pos = lookup(i);
                vals[pos] = __dadd_rn(vals[pos], __dmul_rn(beta, a));
            }
        } else {
            int pos = 0;
            if (active) pos = lookup(i);
            const double prod = __dmul_rn(beta, a);
            unsigned pend = __ballot_sync(0xffffffffu, active);          // duplicates: storage (= lane) order
            while (pend) {
                const int l = __ffs(pend) - 1;
                if (lane == l) vals[pos] = __dadd_rn(vals[pos], prod);
                pend &= pend - 1;
                __syncwarp();
            }
        }
        __syncwarp();
    };

    for (int idx = blockIdx.x * 8 + wid; idx < ncols; idx += nwarps) {
        const int j = list[idx];
        const int nblk = nblk_in[j];
        const int out = Cp[j];
        const int cnt = Cp[j + 1] - out;
        // ---- the block table: insert (distinct blocks), offsets, row list -------------------------
        int my_slot[BLK_STRIDE / 32];
        int running = 0;
#pragma unroll
        for (int r = 0; r < BLK_STRIDE / 32; r++) {
            my_slot[r] = -1;
            if (r * 32 < nblk) {                                  // warp-uniform
                const int t = r * 32 + lane;
                const bool valid = t < nblk;
                int2 bm = make_int2(0, 0);
                if (valid) bm = blk_in[(size_t)j * BLK_STRIDE + t];
                const int c = valid ? __popc((unsigned)bm.y) : 0;
                int inc = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
                const int first = running + inc - c;
                const int word = (bm.x << BLK2_BASE_BITS) | first;
                unsigned h = hash_row(bm.x, LOGH);
                bool pend = valid;
                while (__any_sync(0xffffffffu, pend)) {           // racing plain stores: the warp owns the table
                    if (pend && tab[h].x == EMPTY) tab[h].x = word;
                    __syncwarp();
                    if (pend) {
                        if (tab[h].x == word) { tab[h].y = bm.y; pend = false; }
                        else h = (h + 1) & (H - 1);
                    }
                }
                if (valid) {
                    my_slot[r] = (int)h;
                    unsigned mk = (unsigned)bm.y;
                    int q = first;
                    while (mk) { const int bit = __ffs(mk) - 1; rows[q++] = (bm.x << 5) + bit; mk &= mk - 1; }
                }
                running += __shfl_sync(0xffffffffu, inc, 31);
            }
        }
        // -0.0 is the exact additive identity (-0.0 + x == x bit for bit, also for x == -0.0), so the
        // first product of a row lands as the reference's first-touch assignment
        if (VALUES) for (int t = lane; t < cnt; t += 32) vals[t] = -0.0;
        __syncwarp();
        // ---- accumulate --------------------------------------------------------------------------
        if (VALUES) {
            const int pb_end = Bp[j + 1];
            for (int pb0 = Bp[j]; pb0 < pb_end; pb0 += 32) {
                const int my_pb = pb0 + lane;
                int4 st = make_int4(0, 0, 0, 0);
                if (my_pb < pb_end) {
                    const int k = Bi[my_pb];
                    const int ab = Ap[k], ae = Ap[k + 1];
                    const double beta = Bx[my_pb];
                    st = make_int4(ab, ae - ab, __double2loint(beta), __double2hiint(beta));
                }
                __syncwarp();                                      // the previous chunk has been consumed
                stage[lane] = st;
                __syncwarp();
                const int nb = min(32, pb_end - pb0);
                const bool long_cols = __reduce_max_sync(0xffffffffu, st.y) > 32;   // warp-uniform
                auto run_steps = [&](auto long_tag) {              // two copies: the usual one has no tail loop
                    constexpr bool LONG = decltype(long_tag)::value;
                    for (int s = 0; s < nb; s++) {
                        const int4 g = stage[s];                   // one broadcast LDS.128
                        const double beta = __hiloint2double(g.w, g.z);
                        {
                            int i = 0;
                            double a = 0.0;
                            const bool active = lane < g.y;
                            if (active) { i = Ai[g.x + lane]; a = Ax[g.x + lane]; }
                            accumulate(active, i, a, beta);
                        }
                        if (LONG) {
                            for (int o0 = 32; o0 < g.y; o0 += 32) {    // columns longer than a warp
                                const bool active = o0 + lane < g.y;
                                int i = 0;
                                double a = 0.0;
                                if (active) { i = Ai[g.x + o0 + lane]; a = Ax[g.x + o0 + lane]; }
                                accumulate(active, i, a, beta);
                            }
                        }
                    }
                };
                if (long_cols) run_steps(std::true_type{}); else run_steps(std::false_type{});
            }
        }
        // ---- emit, empty the table -----------------------------------------------------------------
        for (int t = lane; t < cnt; t += 32) {
            Ci[out + t] = rows[t];
            if (VALUES) Cx[out + t] = vals[t];
        }
#pragma unroll
        for (int r = 0; r < BLK_STRIDE / 32; r++) if (my_slot[r] >= 0) tab[my_slot[r]].x = EMPTY;
        __syncwarp();
    }
}

// ---- numeric, one CTA per column, dense workspaces in global memory -----------------
// The reference's algorithm verbatim per column: marks w[m], accumulator x[m]
// (csparse.py:1624-1626), discovery order by a block-wide rank of the new rows.
template <bool VALUES, bool CANON>
__global__ void __launch_bounds__(256)
k_num_dense(const int *__restrict__ list, int ncols, int m,
            const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
            const csi *__restrict__ Bp, const csi *__restrict__ Bi, const double *__restrict__ Bx,
            const csi *__restrict__ Cp, csi *Ci, double *Cx, int *marks, double *acc)
{
    __shared__ int s_warp[8];
    __shared__ int s_cnt;
    int *w = marks + (size_t)blockIdx.x * m;
    double *x = VALUES ? acc + (size_t)blockIdx.x * m : nullptr;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt = lanemask_lt();
    int mark = 0;
    for (int idx = blockIdx.x; idx < ncols; idx += gridDim.x) {
        const int j = list[idx];
        mark++;
        const int out = Cp[j];
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        for (int pb = Bp[j]; pb < Bp[j + 1]; pb++) {
            const int k = Bi[pb];
            const double beta = VALUES ? Bx[pb] : 1.0;
            const int ab = Ap[k], ae = Ap[k + 1];
            for (int pa0 = ab; pa0 < ae; pa0 += 256) {
                const int pa = pa0 + threadIdx.x;
                const bool active = pa < ae;
                int i = 0;
                double prod = 0.0;
                bool isnew = false;
                if (active) {
                    i = Ai[pa];
                    if (VALUES) prod = __dmul_rn(beta, Ax[pa]);
                    if (CANON) { isnew = w[i] < mark; if (isnew) w[i] = mark; }
                    else       { isnew = atomicMax(&w[i], mark) < mark; }
                }
                const unsigned nm = __ballot_sync(0xffffffffu, isnew);
                if (lane == 0) s_warp[wid] = __popc(nm);
                __syncthreads();
                int before = s_cnt;
                for (int u = 0; u < wid; u++) before += s_warp[u];
                if (isnew) {
                    Ci[out + before + __popc(nm & lt)] = i;
                    if (VALUES) x[i] = CANON ? prod : 0.0;
                }
                __syncthreads();
                if (threadIdx.x == 0) {
                    int tot = 0;
                    for (int u = 0; u < 8; u++) tot += s_warp[u];
                    s_cnt += tot;
                }
                if (VALUES && active) {
                    if (CANON) { if (!isnew) x[i] = __dadd_rn(x[i], prod); }
                    else       atomicAdd(&x[i], prod);
                }
                __syncthreads();
            }
        }
        if (VALUES) {
            const int cnt = s_cnt;
            for (int t = threadIdx.x; t < cnt; t += 256) Cx[out + t] = x[Ci[out + t]];   // csparse.py:1637-1639
        }
        __syncthreads();
    }
}

// ---- "does every column have strictly increasing rows?" -------------------------------
__global__ void k_canon(const csi *__restrict__ Ap, const csi *__restrict__ Ai, int n, int nnz, int *bad)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x + 1;
    if (p >= nnz) return;
    if (Ai[p - 1] >= Ai[p]) {
        const int j = upper_row(Ap, 0, n, p);      // column holding entry p
        if (Ap[j] != p) *bad = 1;                  // descent strictly inside a column
    }
}

// exact column sizes for the numeric classes; columns with a template (cb[j] >= 0) are k_num_tpl's
__global__ void k_col_sizes(int n, const csi *__restrict__ Cp, int *__restrict__ sz, const int *__restrict__ cb)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) sz[j] = (cb && cb[j] >= 0) ? 0 : Cp[j + 1] - Cp[j];
}

int mat_is_canonical(csb200_mat *A, int *out)
{
    if (A->canon < 0) {
        if (A->nnz < 2) { A->canon = 1; }
        else {
            DevBuf<int> bad;
            CSB_TRY(bad.alloc(1));
            CSB_CUDA(cudaMemsetAsync(bad.ptr, 0, sizeof(int), stream()));
            k_canon<<<ceil_div(A->nnz, 256), 256, 0, stream()>>>(A->p, A->i, A->n, (int)A->nnz, bad.ptr);
            CSB_LAUNCHED();
            int h = 0;
            CSB_CUDA(cudaMemcpyAsync(&h, bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream()));
            CSB_CUDA(cudaStreamSynchronize(stream()));
            A->canon = h ? 0 : 1;
        }
    }
    *out = A->canon;
    return CSB200_OK;
}

// ---- host side ---------------------------------------------------------------------------
// pattern classes of A's columns (spgemm_tpl.cuh), once per handle: A->cls[k] = class or -1
static int ensure_classes(csb200_mat *A)
{
    if (A->cls_state >= 0) return CSB200_OK;
    A->cls_state = 0;
    if (A->n == 0 || A->nnz == 0) return CSB200_OK;
    cudaStream_t s = stream();
    DevBuf<unsigned char> tb;
    CSB_TRY(tb.alloc(CLS_TABLE_BYTES));
    int *ca = nullptr;
    CSB_TRY(dev_alloc(&ca, (size_t)A->n));
    const ClsTable t = cls_table_at(tb.ptr);
    const int grid = ceil_div(A->n, 256);
    k_cls_init<<<CLS_SLOTS / 256, 256, 0, s>>>(t);
    k_cls_hash_a<<<grid, 256, 0, s>>>(A->n, A->p, A->i, t, ca);
    k_cls_compact<<<1, 1024, 0, s>>>(t);
    k_cls_verify_a<<<grid, 256, 0, s>>>(A->n, A->p, A->i, t, ca);
    g_launches.fetch_add(4, std::memory_order_relaxed);
    int info[4] = {0, 1, 0, 0};
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(info, t.info, sizeof(info), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { dev_free(ca); return set_error(CSB200_ERR_CUDA, "pattern classes: %s", cudaGetErrorString(e)); }
    if (info[1] || info[2] == 0) { dev_free(ca); return CSB200_OK; }
    A->cls = ca;
    A->cls_count = info[2];
    A->cls_state = 1;
    return CSB200_OK;
}

// entry-major copy of the values (spgemm_tpl.cuh, k_num_soa), once per handle; soa_state 0 = the
// matrix does not qualify (a column longer than SOA_MAXLEN, or a copy more than twice the matrix)
static int ensure_soa(csb200_mat *M)
{
    if (M->soa_state >= 0) return CSB200_OK;
    M->soa_state = 0;
    if (!M->x || M->n == 0 || M->nnz == 0) return CSB200_OK;
    cudaStream_t s = stream();
    int h_max = 0;
    {
        DevBuf<int> d_max;
        CSB_TRY(d_max.alloc(1));
        CSB_CUDA(cudaMemsetAsync(d_max.ptr, 0, sizeof(int), s));
        k_max_col_len<<<min(ceil_div(M->n, 256), sm_count() * 8), 256, 0, s>>>(M->n, M->p, d_max.ptr);
        CSB_LAUNCHED();
        CSB_CUDA(cudaMemcpyAsync(&h_max, d_max.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
        CSB_CUDA(cudaStreamSynchronize(s));
    }
    M->max_col_len = h_max;
    const long long slots = (long long)h_max * M->n;
    if (h_max == 0 || h_max > SOA_MAXLEN || slots > 2 * M->nnz + 1024 || slots >= (1LL << 30)) return CSB200_OK;
    CSB_TRY(dev_alloc(&M->soa_x, (size_t)slots));
    k_soa_build<<<ceil_div(M->n, 256), 256, 0, s>>>(M->n, h_max, M->p, M->x, M->soa_x);
    CSB_LAUNCHED();
    M->soa_len = h_max;
    M->soa_state = 1;
    return CSB200_OK;
}

// ncols_dev != null: the number of listed columns is read on the device (no host round trip)
static int ensure_compressed(csb200_mat *A)
{
    if (A->c32_len) return CSB200_OK;
    const size_t cap = (size_t)(A->nnz > 0 ? A->nnz : 1);
    CSB_TRY(dev_alloc(&A->c32_blk, cap));
    CSB_TRY(dev_alloc(&A->c32_mask, cap));
    csi *len = nullptr;
    CSB_TRY(dev_alloc(&len, (size_t)A->n + 1));
    if (A->n > 0) {
        k_compress_cols<<<ceil_div((long long)A->n * 32, 256), 256, 0, stream()>>>(A->n, A->p, A->i, A->c32_blk,
                                                                                 A->c32_mask, len);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { dev_free(len); return set_error(CSB200_ERR_CUDA, "k_compress_cols: %s", cudaGetErrorString(e)); }
    }
    A->c32_len = len;
    return CSB200_OK;
}

template <int LOGH, int CAP, int WARPS, bool OPTIMISTIC>
static int run_sym_flat(const int *list, const int *ncols_dev, int ncols, const csb200_mat *A,
                        const csb200_mat *B, int *cnt, int *ovf_list, int *ovf_count,
                        int2 *blk_out = nullptr, int *nblk_out = nullptr)
{
    if (!ncols_dev && ncols == 0) return CSB200_OK;
    const size_t smem = (size_t)WARPS * ((size_t)(1 << LOGH) * 8 + CAP * 2);
    const int resident = (int)min((size_t)(2048 / (WARPS * 32)), (size_t)(220 * 1024) / smem);
    const int grid = ncols_dev ? sm_count() * resident
                               : (int)min((long long)ceil_div(ncols, WARPS), (long long)sm_count() * resident);
    auto kern = k_sym_flat<LOGH, CAP, WARPS, OPTIMISTIC>;
    CSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, WARPS * 32, smem, stream()>>>(list, ncols_dev, ncols, A->p, A->c32_blk, A->c32_mask, A->c32_len,
                                               B->p, B->i, cnt, ovf_list, ovf_count, blk_out, nblk_out);
    CSB_LAUNCHED();
    return CSB200_OK;
}

template <int LOGH, int CAP, int WARPS>
static int run_num_warp(const int *list, int ncols, const csb200_mat *A, const csb200_mat *B,
                        csb200_mat *C, bool values, bool canon)
{
    if (ncols == 0) return CSB200_OK;
    const size_t smem = (size_t)WARPS * ((size_t)(1 << LOGH) * 12 + CAP * 2);
    const int resident = (int)min((size_t)16, (size_t)(220 * 1024) / smem);
    const int grid = (int)min((long long)ceil_div(ncols, WARPS), (long long)sm_count() * resident);
#define NUM_LAUNCH(V, K)                                                                          \
    do {                                                                                          \
        auto kern = k_num_warp<LOGH, CAP, WARPS, V, K>;                                           \
        CSB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, WARPS * 32, smem, stream()>>>(list, ncols, A->p, A->i, A->x, B->p, B->i, B->x, \
                                                   C->p, C->i, C->x);                             \
    } while (0)
    if (values) { if (canon) NUM_LAUNCH(true, true); else NUM_LAUNCH(true, false); }
    else        { if (canon) NUM_LAUNCH(false, true); else NUM_LAUNCH(false, false); }
#undef NUM_LAUNCH
    CSB_LAUNCHED();
    return CSB200_OK;
}

// numeric class selection for the blocked kernel: 1 where the symbolic phase kept the column's
// block list and the column fits BLK_CAP rows; the ordered kernels then skip it (size 0)
__global__ void k_pick_blocked(int n, const int *__restrict__ nblk, int *__restrict__ size, int *__restrict__ pick)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const bool b = nblk[j] >= 0 && size[j] > 0 && size[j] <= BLK_CAP;
    pick[j] = b ? 1 : 0;
    if (b) size[j] = 0;
}


// ordered: the columns of C must come out in the reference's discovery order (cs_add / cs_dupl are
// built on that); otherwise the blocked numeric kernel may emit them block by block.
int multiply_impl(csb200_mat *A, csb200_mat *B, csb200_mat **out, bool ordered)
{
    const csi m = A->m, n = B->n;
    if (tls().multiply_ordered) ordered = true;
    const bool values = A->x != nullptr && B->x != nullptr;      // csparse.py:1625
    cudaStream_t s = stream();

    csb200_mat *C = new csb200_mat();
    C->m = m; C->n = n; C->device = A->device;
    auto fail = [&](int st) { csb200_mat_free(C); return st; };
    int st = dev_alloc(&C->p, (size_t)n + 1 + MAT_PAD);
    if (st != CSB200_OK) return fail(st);

    int canon = 1;
    if ((st = mat_is_canonical(A, &canon)) != CSB200_OK) return fail(st);

    // pattern classes (spgemm_tpl.cuh): 0 automatic, 1 off, 2 also below TPL_MIN_N columns (tests)
    const int tmode = tls().multiply_templates;
    bool tpl = n > 0 && canon && tmode != 1 && A->nnz > 0 && B->nnz > 0 &&
               (tmode >= 2 || (n >= TPL_MIN_N && A->n >= TPL_MIN_N));
    if (tpl) {
        if ((st = ensure_classes(A)) != CSB200_OK) return fail(st);
        tpl = A->cls_state == 1;
    }
    tls().last_templated = 0;
    // lane-per-column numeric kernel on entry-major copies of both factors' values
    bool soa = tpl && values && tls().multiply_templates != 3;
    if (soa) {
        if ((st = ensure_soa(A)) != CSB200_OK || (st = ensure_soa(B)) != CSB200_OK) return fail(st);
        soa = A->soa_state == 1 && B->soa_state == 1;
    }

    arena_hint((size_t)(n > 0 ? n : 1) * (9 * 4 + BLK_STRIDE * sizeof(int2)) + (1 << 16) +
               (tpl ? CLS_TABLE_BYTES + (size_t)CLS_MAX * (4 + TPL_UB + TPL_CAP * 4) + 5 * (size_t)(n > 0 ? n : 1) : 0) +
               (soa ? (size_t)CLS_MAX * (1 + SOA_TERMS * (sizeof(int2) + 1) + (SOA_CHUNKS * SOA_WARPS + 1) * 4) : 0));
    DevBuf<int> ub, cnt, lists, counts, marks, marks_num, nblk, pick, cb, tpl_cnt, tpl_rows, tpl_wptr;
    DevBuf<unsigned char> tpl_table, tpl_pos, tpl_soa, tpl_mode;
    DevBuf<int> left_list;
    DevBuf<unsigned char> tpl_tend;
    DevBuf<int> tpl_terms;
    DevBuf<int2> blkbuf;
    DevBuf<double> acc;
    // the blocked numeric path keeps BLK_STRIDE pairs per column (512 B): only while that stays modest
    const bool blocked = !ordered && n > 0 && (size_t)n * BLK_STRIDE * sizeof(int2) <= ((size_t)4 << 30);
    DevBuf<unsigned long long> flops;
    DevBuf<long long> total;
    const size_t ncap = (size_t)(n > 0 ? n : 1);
    if ((st = ub.alloc(ncap)) || (st = cnt.alloc(ncap)) || (st = lists.alloc(4 * ncap)) ||
        (st = counts.alloc(8)) || (st = flops.alloc(1)) || (st = total.alloc(1)))
        return fail(st);
#define MM_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
        set_error(CSB200_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); return fail(CSB200_ERR_CUDA); } } while (0)
#define MM_LAUNCHED() do { g_launches.fetch_add(1, std::memory_order_relaxed); MM_CUDA(cudaGetLastError()); } while (0)
#define MM_TRY(expr) do { int s_ = (expr); if (s_ != CSB200_OK) return fail(s_); } while (0)
    MM_CUDA(cudaMemsetAsync(counts.ptr, 0, 8 * sizeof(int), s));
    MM_CUDA(cudaMemsetAsync(flops.ptr, 0, sizeof(unsigned long long), s));
    MM_CUDA(cudaMemsetAsync(cnt.ptr, 0, ncap * sizeof(int), s));

    int h_counts[8] = {0};
    bool generic_cols = true;
    int h_tpl[8] = {0, 0, 0, 0, 0, 0, 0, 0};          // ClsTable::info: [3] templated columns, [4] of them for k_num_tpl
    unsigned long long h_flops = 0;
    constexpr int DENSE_CTAS = 64;
    ClsTable tb{};
    if (n > 0) {
        if (!tpl) {
            k_ub<<<ceil_div(n, 256), 256, 0, s>>>(n, B->p, B->i, A->p, ub.ptr, flops.ptr);
            MM_LAUNCHED();
        }
        if (tpl) {
            // classes of B's columns, one template per class, cnt[] of every column that has one
            MM_TRY(tpl_table.alloc(CLS_TABLE_BYTES));
            MM_TRY(cb.alloc(ncap));
            MM_TRY(tpl_cnt.alloc(CLS_MAX));
            MM_TRY(tpl_pos.alloc((size_t)CLS_MAX * TPL_UB));
            MM_TRY(tpl_rows.alloc((size_t)CLS_MAX * TPL_CAP));
            MM_TRY(tpl_mode.alloc(ncap));
            MM_TRY(left_list.alloc(ncap));
            if (soa) {
                MM_TRY(tpl_soa.alloc(CLS_MAX));
                MM_TRY(tpl_terms.alloc((size_t)CLS_MAX * SOA_TERMS));
                MM_TRY(tpl_tend.alloc((size_t)CLS_MAX * SOA_TERMS));
                MM_TRY(tpl_wptr.alloc((size_t)CLS_MAX * (SOA_CHUNKS * SOA_WARPS + 1)));
            }
            tb = cls_table_at(tpl_table.ptr);
            const int wgrid = ceil_div(n, 256);
            k_cls_init<<<CLS_SLOTS / 256, 256, 0, s>>>(tb);
            MM_LAUNCHED();
            k_cls_hash_b<<<wgrid, 256, 0, s>>>(n, B->p, B->i, A->p, A->cls, tb, cb.ptr, ub.ptr, flops.ptr);    // k_ub's work included
            MM_LAUNCHED();
            k_cls_compact<<<1, 1024, 0, s>>>(tb);
            MM_LAUNCHED();
            k_tpl_build<<<CLS_MAX, 32, 0, s>>>(tb, A->p, A->i, B->p, B->i, ub.ptr, tpl_cnt.ptr, tpl_pos.ptr, tpl_rows.ptr,
                                               A->n, n, soa ? A->soa_len : 0, soa ? B->soa_len : 0,
                                               soa ? tpl_soa.ptr : nullptr, tpl_terms.ptr, tpl_tend.ptr, tpl_wptr.ptr);
            MM_LAUNCHED();
            k_cls_verify_apply<<<wgrid, 256, 0, s>>>(n, B->p, B->i, A->cls, tb, tpl_cnt.ptr, cb.ptr, cnt.ptr, ub.ptr,
                                                     soa ? tpl_soa.ptr : nullptr, tpl_mode.ptr, left_list.ptr);
            MM_LAUNCHED();
            MM_CUDA(cudaMemcpyAsync(h_tpl, tb.info, sizeof(h_tpl), cudaMemcpyDeviceToHost, s));
        }
        // symbolic classes by min(ub, m): <=256 (cannot outgrow the small table) |
        // <=8000 (optimistic: small table first, columns that outgrow it -> list 2) | dense
        k_bin<<<ceil_div(n, 256), 256, 0, s>>>(n, ub.ptr, m, 256, 8000, 8000, lists.ptr, counts.ptr);
        MM_LAUNCHED();
        MM_CUDA(cudaMemcpyAsync(h_counts, counts.ptr, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
        MM_CUDA(cudaMemcpyAsync(&h_flops, flops.ptr, sizeof(h_flops), cudaMemcpyDeviceToHost, s));
        MM_CUDA(cudaStreamSynchronize(s));
        tls().last_flops = (int64_t)h_flops;
        tls().last_templated = h_tpl[3];
        int *ovf_list = lists.ptr + 2 * ncap, *ovf_count = counts.ptr + 2;
        generic_cols = h_counts[0] + h_counts[1] + h_counts[2] + h_counts[3] > 0;     // columns without a template
        if (h_counts[0] + h_counts[1] > 0) MM_TRY(ensure_compressed(A));
        if (blocked) {
            MM_TRY(nblk.alloc(ncap));
            MM_TRY(pick.alloc(ncap));
            MM_TRY(blkbuf.alloc((size_t)n * BLK_STRIDE));
            MM_CUDA(cudaMemsetAsync(nblk.ptr, 0xff, ncap * sizeof(int), s));
        }
        MM_TRY((run_sym_flat<9, 256, 8, false>(lists.ptr, nullptr, h_counts[0], A, B, cnt.ptr, nullptr, nullptr,
                                               blkbuf.ptr, nblk.ptr)));
        if (h_counts[1] > 0) {
            MM_TRY((run_sym_flat<9, 256, 8, true>(lists.ptr + ncap, nullptr, h_counts[1], A, B, cnt.ptr,
                                                  ovf_list, ovf_count, blkbuf.ptr, nblk.ptr)));
            MM_TRY((run_sym_flat<13, 8000, 2, false>(ovf_list, ovf_count, 0, A, B, cnt.ptr, nullptr, nullptr)));
        }
        if (h_counts[3] > 0) {
            const int ctas = min(DENSE_CTAS, h_counts[3]);
            MM_TRY(marks.alloc((size_t)ctas * m));
            MM_CUDA(cudaMemsetAsync(marks.ptr, 0, (size_t)ctas * m * sizeof(int), s));
            k_sym_dense<<<ctas, 256, 0, s>>>(lists.ptr + 3 * ncap, h_counts[3], m, A->p, A->i, B->p, B->i,
                                              marks.ptr, cnt.ptr);
            MM_LAUNCHED();
        }
    } else {
        tls().last_flops = 0;
    }
    // Cp = cumsum(cnt); nnz(C)
    MM_TRY(launch_excl_scan(C->p, cnt.ptr, n, total.ptr, nullptr));
    // cnt now holds Cp[0..n-1]; the class of a column for the numeric phase is
    // decided by its exact size Cp[j+1]-Cp[j], recomputed into ub.
    long long h_total = 0;
    MM_CUDA(cudaMemcpyAsync(&h_total, total.ptr, sizeof(long long), cudaMemcpyDeviceToHost, s));
    MM_CUDA(cudaStreamSynchronize(s));
    if (h_total > 0x7fffffffLL) {
        set_error(CSB200_ERR_OVERFLOW, "cs_multiply: nnz(C) = %lld does not fit int32", h_total);
        return fail(CSB200_ERR_OVERFLOW);
    }
    C->nnz = h_total;
    const size_t cap = (size_t)(h_total > 0 ? h_total : 1) + MAT_PAD;
    MM_TRY(dev_alloc(&C->i, cap));
    if (values) MM_TRY(dev_alloc(&C->x, cap));
    if (h_total > 0) {
        // exact sizes: ub[j] = Cp[j+1] - Cp[j]
        k_col_sizes<<<ceil_div(n, 256), 256, 0, s>>>(n, C->p, ub.ptr, tpl ? cb.ptr : nullptr);
        MM_LAUNCHED();
        if (h_tpl[3] > 0 && soa) {
            // 32 columns per CTA in lock step; what it leaves (grid boundaries, long columns) goes to k_num_tpl
            // few CTAs per SM: each works on ~100 KB of A that should stay in its SM's L1
            static const int soa_ctas = getenv("CSB200_SOA_CTAS") ? atoi(getenv("CSB200_SOA_CTAS")) : 4;
            const int grid = (int)min((long long)ceil_div(n, 32), (long long)sm_count() * soa_ctas);
            static const int soa_stride = getenv("CSB200_SOA_STRIDE") ? atoi(getenv("CSB200_SOA_STRIDE")) : 2;   // blocks per ticket; 0: round robin (A/B)
            DevBuf<int> soa_tickets;
            MM_TRY(soa_tickets.alloc((size_t)ceil_div(n, 32) + 1));
            if (soa_stride > 0) MM_CUDA(cudaMemsetAsync(soa_tickets.ptr, 0, ((size_t)ceil_div(n, 32) + 1) * sizeof(int), s));
#define SOA_LAUNCH(MINB)                                                                                       \
            do {                                                                                               \
                MM_CUDA(cudaFuncSetAttribute(k_num_soa<MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SOA_SMEM)); \
                k_num_soa<MINB><<<grid, SOA_THREADS, SOA_SMEM, s>>>(n, cb.ptr, tpl_cnt.ptr, tpl_mode.ptr, tpl_terms.ptr, tpl_tend.ptr, \
                                                                  tpl_wptr.ptr, tpl_rows.ptr, A->soa_x, B->soa_x, A->n, n, C->p, C->i, C->x, soa_stride, soa_tickets.ptr); \
            } while (0)
            if (soa_ctas >= 6) SOA_LAUNCH(6); else if (soa_ctas == 5) SOA_LAUNCH(5); else SOA_LAUNCH(4);
#undef SOA_LAUNCH
            MM_LAUNCHED();
        }
        if (h_tpl[4] > 0) {
            constexpr int smem = 8 * TPL_PER_WARP;
            const int n_left = h_tpl[4];
            static const int variant = getenv("CSB200_TPL_VARIANT") ? atoi(getenv("CSB200_TPL_VARIANT")) : 0;
#define TPL_LAUNCH(V, BATCH, MINB)                                                                           \
            do {                                                                                             \
                auto kern = k_num_tpl<V, BATCH, MINB>;                                                       \
                const int grid = (int)min((long long)ceil_div(n_left, 8), (long long)sm_count() * MINB);     \
                MM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));      \
                kern<<<grid, 256, smem, s>>>(left_list.ptr, tb.info + 4, cb.ptr, tpl_cnt.ptr, tpl_pos.ptr, tpl_rows.ptr, A->p, A->x,  \
                                             B->p, B->i, B->x, C->p, C->i, C->x);                            \
            } while (0)
            if (!values) TPL_LAUNCH(false, 4, 8);
            else if (variant == 1) TPL_LAUNCH(true, 8, 4);
            else if (variant == 2) TPL_LAUNCH(true, 8, 3);
            else TPL_LAUNCH(true, 4, 5);                 // fastest of the three on st27 128^3 (5.36 / 5.77 / 5.60 ms per product)
#undef TPL_LAUNCH
            MM_LAUNCHED();
        }
        int n_blocked = 0;
        if (blocked && generic_cols) {
            // columns whose block list was kept go to the blocked kernel (list 3 region is free here)
            k_pick_blocked<<<ceil_div(n, 256), 256, 0, s>>>(n, nblk.ptr, ub.ptr, pick.ptr);
            MM_LAUNCHED();
            MM_CUDA(cudaMemsetAsync(counts.ptr, 0, 8 * sizeof(int), s));
            k_bin<<<ceil_div(n, 256), 256, 0, s>>>(n, pick.ptr, INT_MAX, 1, 1, 1, lists.ptr, counts.ptr);
            MM_LAUNCHED();
            MM_CUDA(cudaMemcpyAsync(&n_blocked, counts.ptr, sizeof(int), cudaMemcpyDeviceToHost, s));
            MM_CUDA(cudaStreamSynchronize(s));
            // kernel version: 2 = the first blocked kernel, 3 = k_num_blocked2 (csb200_multiply_force_path);
            // the packed table slot of the second needs rows < 2^29
            int version = tls().multiply_blocked_version ? tls().multiply_blocked_version : 3;
            if ((long long)m > BLK2_MAX_ROWS) version = 2;
            if (n_blocked > 0 && version >= 3) {
                constexpr int smem = 8 * BLK2_PER_WARP;
                const int grid = (int)min((long long)ceil_div(n_blocked, 8), (long long)sm_count() * 8);
#define BLK2_LAUNCH(V, K)                                                                             \
                do {                                                                                      \
                    auto kern = k_num_blocked2<V, K>;                                                     \
                    MM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
                    kern<<<grid, 256, smem, s>>>(lists.ptr, n_blocked, A->p, A->i, A->x, B->p, B->i, B->x,   \
                                                 blkbuf.ptr, nblk.ptr, C->p, C->i, C->x);                 \
                } while (0)
                if (values) { if (canon) BLK2_LAUNCH(true, true); else BLK2_LAUNCH(true, false); }
                else        { if (canon) BLK2_LAUNCH(false, true); else BLK2_LAUNCH(false, false); }
#undef BLK2_LAUNCH
                MM_LAUNCHED();
            } else if (n_blocked > 0) {
                constexpr int smem = 8 * (BLK_CAP * 12 + BLK_H * 12);
                const int grid = (int)min((long long)ceil_div(n_blocked, 8), (long long)sm_count() * min(8, (220 * 1024) / smem));
#define BLK_LAUNCH(V, K)                                                                              \
                do {                                                                                      \
                    auto kern = k_num_blocked<V, K>;                                                      \
                    MM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
                    kern<<<grid, 256, smem, s>>>(lists.ptr, n_blocked, A->p, A->i, A->x, B->p, B->i, B->x,   \
                                                 blkbuf.ptr, nblk.ptr, C->p, C->i, C->x);                 \
                } while (0)
                if (values) { if (canon) BLK_LAUNCH(true, true); else BLK_LAUNCH(true, false); }
                else        { if (canon) BLK_LAUNCH(false, true); else BLK_LAUNCH(false, false); }
#undef BLK_LAUNCH
                MM_LAUNCHED();
            }
        }
        // numeric classes by exact column size: <=128 | <=512 | <=2048 | dense (nothing to do, and no
        // host round trip, when every column came from a template)
        h_counts[0] = h_counts[1] = h_counts[2] = h_counts[3] = 0;
        if (generic_cols) {
            MM_CUDA(cudaMemsetAsync(counts.ptr, 0, 8 * sizeof(int), s));
            k_bin<<<ceil_div(n, 256), 256, 0, s>>>(n, ub.ptr, INT_MAX, 128, 512, 2048, lists.ptr, counts.ptr);
            MM_LAUNCHED();
            MM_CUDA(cudaMemcpyAsync(h_counts, counts.ptr, 4 * sizeof(int), cudaMemcpyDeviceToHost, s));
            MM_CUDA(cudaStreamSynchronize(s));
        }
        MM_TRY((run_num_warp<8, 128, 8>(lists.ptr, h_counts[0], A, B, C, values, canon != 0)));
        MM_TRY((run_num_warp<10, 512, 8>(lists.ptr + ncap, h_counts[1], A, B, C, values, canon != 0)));
        MM_TRY((run_num_warp<12, 2048, 4>(lists.ptr + 2 * ncap, h_counts[2], A, B, C, values, canon != 0)));
        if (h_counts[3] > 0) {
            const int ctas = min(DENSE_CTAS, h_counts[3]);
            MM_TRY(marks_num.alloc((size_t)ctas * m));
            MM_CUDA(cudaMemsetAsync(marks_num.ptr, 0, (size_t)ctas * m * sizeof(int), s));
            if (values) MM_TRY(acc.alloc((size_t)ctas * m));
            const int *lst = lists.ptr + 3 * ncap;
#define DENSE_LAUNCH(V, K) k_num_dense<V, K><<<ctas, 256, 0, s>>>(lst, h_counts[3], m, A->p, A->i, A->x, \
                                B->p, B->i, B->x, C->p, C->i, C->x, marks_num.ptr, acc.ptr)
            if (values) { if (canon) DENSE_LAUNCH(true, true); else DENSE_LAUNCH(true, false); }
            else        { if (canon) DENSE_LAUNCH(false, true); else DENSE_LAUNCH(false, false); }
#undef DENSE_LAUNCH
            MM_LAUNCHED();
        }
    }
#undef MM_CUDA
#undef MM_LAUNCHED
#undef MM_TRY
    *out = C;
    return CSB200_OK;
}

}  // namespace csb
