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
constexpr int SOA_UB = 1024;             // products per column the lane-per-column kernel holds (term table in shared memory)
constexpr int SOA_CNT = 128;             // rows per column it holds
constexpr int SOA_MAXLEN = 64;           // longest column of an entry-major copy
constexpr int SOA_MIN_LANES = 8;         // columns of one class inside a 32-column block worth a pass
constexpr int SOA_THREADS = 256;
constexpr int SOA_WARPS = SOA_THREADS / 32;
constexpr int SOA_CHUNKS = SOA_CNT / 32;
constexpr int SOA_BATCH = 8;             // terms a warp takes per step (its lists are padded to a multiple)
constexpr int SOA_TERMS = SOA_UB + SOA_CHUNKS * SOA_WARPS * (SOA_BATCH - 1);
// A term of k_num_soa in one 32-bit word: e (entry of the A column, 6 bits) | d (entry of B(:,j), 6 bits) |
// k - j (signed, 20 bits).  The term lists are read by all lanes of a warp at once, and a uniform
// shared-memory load costs a wavefront per 4 bytes like any other: {A offset, B offset} pairs of 8 bytes
// were a third of the kernel's shared-memory + L1 wavefronts (profiles/r2_notes.md).
constexpr int SOA_DK_LIM = 1 << 19;
__host__ __device__ __forceinline__ int soa_term(int e, int d, int dk) { return (e & 63) | ((d & 63) << 6) | (dk << 12); }
static_assert(SOA_MAXLEN <= 64, "six bits per entry number");

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

// ---- classes of the columns: one THREAD per column -----------------------------------------------
// A thread walks its column (the strided reads of neighbouring threads share lines through L1, as in
// k_ub) and hashes it; inside a warp a column that repeats its left neighbour's hash takes that
// neighbour's class, so only the heads of runs go to the table -- two million columns of one class
// would otherwise queue on a single L2 sector.  Columns longer than CLS_MAXLEN have no class.
constexpr int CLS_MAXLEN = 1024;

// slot of every lane's class from the lanes' hashes (ok = the lane has one); all 32 lanes call
__device__ __forceinline__ int cls_resolve(const ClsTable &t, unsigned long long h, bool ok, int col)
{
    const int lane = threadIdx.x & 31;
    const unsigned long long hp = __shfl_up_sync(0xffffffffu, h, 1);
    const bool okp = __shfl_up_sync(0xffffffffu, (int)ok, 1) != 0;
    const bool head = ok && (lane == 0 || !okp || hp != h);
    int slot = -1;
    if (head) slot = cls_insert(t, h, col);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    const unsigned left = heads & (0xffffffffu >> (31 - lane));      // heads at or left of this lane
    const int src = left ? 31 - __clz(left) : lane;
    slot = __shfl_sync(0xffffffffu, slot, src);
    return ok ? slot : -1;
}

// A: the rows relative to the column index, in storage order
__global__ void __launch_bounds__(256)
k_cls_hash_a(int n, const csi *__restrict__ Ap, const csi *__restrict__ Ai, ClsTable t, int *__restrict__ ca)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h = 0;
    bool ok = false;
    if (k < n && !__ldcg(t.info + 1)) {
        const int b = Ap[k], e = Ap[k + 1];
        ok = e - b <= CLS_MAXLEN;
        if (ok)
            for (int p = b; p < e; p++)
                h += mix64(mix64((unsigned long long)(p - b) + 1) ^ (unsigned long long)(unsigned)(Ai[p] - k));
        h += mix64(0xA5A5A5A5ull + (unsigned long long)(e - b));
        if (h == 0ull) h = 1ull;
    }
    const int slot = cls_resolve(t, h, ok, k);
    if (k < n) ca[k] = slot;
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
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n || t.info[1]) return;                 // the flag is final by now: a cached load
    const int slot = ca[k];
    bool ok = slot >= 0;
    if (ok) {
        const int r = t.rep[slot];
        if (r != k) {
            const int b = Ap[k], len = Ap[k + 1] - b, br = Ap[r];
            ok = (Ap[r + 1] - br) == len;
            for (int e = 0; ok && e < len; e++) ok = Ai[b + e] - k == Ai[br + e] - r;
        }
    }
    ca[k] = ok ? t.dense[slot] : -1;
}

// B: (row relative to the column, class of that column of A); also k_ub's work on the same walk:
// ub[j] = multiply-adds of column j, their total in *flops
__global__ void __launch_bounds__(256)
k_cls_hash_b(int n, const csi *__restrict__ Bp, const csi *__restrict__ Bi, const csi *__restrict__ Ap,
             const int *__restrict__ ca, ClsTable t, int *__restrict__ cb, int *__restrict__ ub,
             unsigned long long *flops)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long h = 0;
    long long s = 0;
    bool ok = false;
    if (j < n) {
        const int b = Bp[j], e = Bp[j + 1];
        ok = e > b && e - b <= CLS_MAXLEN && !__ldcg(t.info + 1);
        for (int p = b; p < e; p++) {
            const int k = Bi[p];
            s += Ap[k + 1] - Ap[k];
            if (ok) {
                const int c = ca[k];
                if (c < 0) ok = false;
                // order-dependent polynomial hash, one 64-bit multiply-add per entry (two mix64 per entry made
                // this kernel issue-bound: 75 instructions per entry); a collision only costs the comparison
                // with the class representative that follows anyway
                h = h * 0x9E3779B97F4A7C15ull + (((unsigned long long)(unsigned)c << 32) | (unsigned long long)(unsigned)(k - j));
            }
        }
        h = mix64(h + mix64(0x5A5A5A5Aull + (unsigned long long)(e - b)));
        if (h == 0ull) h = 1ull;
        ub[j] = (int)min(s, (long long)INT_MAX);
    }
    const int slot = cls_resolve(t, h, ok, j);
    if (j < n) cb[j] = slot;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(flops, (unsigned long long)s);
}

// ---- the template of a class: cs_scatter on its representative column, one warp ------------------
// A's columns are canonical (distinct rows per column), so the rows of a 32-entry step are distinct.
// For k_num_soa the same trace is also stored inverted: for every row t of the column the products
// that land on it, in the reference's order, as offsets into the entry-major copies of A and B
// (term = one word, soa_term(e, d, k - j), for the e-th entry of the A column named by the d-th entry of
// B(:,j)); tpl_soa[c] says whether the class fits that kernel.
__global__ void __launch_bounds__(32)
k_tpl_build(ClsTable t, const csi *__restrict__ Ap, const csi *__restrict__ Ai, const csi *__restrict__ Bp,
            const csi *__restrict__ Bi, const int *__restrict__ ub, int *__restrict__ tpl_cnt,
            unsigned char *__restrict__ tpl_pos, int *__restrict__ tpl_rows,
            int nA, int nB, int soa_len_a, int soa_len_b, unsigned char *__restrict__ tpl_soa,
            int *__restrict__ tpl_terms, unsigned char *__restrict__ tpl_tend, int *__restrict__ tpl_wptr)
{
    constexpr int LOGH = 10, H = 1 << LOGH;
    static_assert(H >= 2 * (TPL_CAP + 32), "the table never fills");
    __shared__ int keys[H];
    __shared__ unsigned short posof[H];
    __shared__ int tcount[TPL_CAP + 1];
    const int lane = threadIdx.x, c = blockIdx.x;
    if (__ldcg(t.info + 1) || c >= __ldcg(t.info + 2)) return;
    const int j = t.rep_dense[c];
    if (lane == 0 && tpl_soa) tpl_soa[c] = 0;
    if (ub[j] > TPL_UB) { if (lane == 0) tpl_cnt[c] = -1; return; }
    for (int s = lane; s < H; s += 32) keys[s] = EMPTY;
    for (int s = lane; s <= TPL_CAP; s += 32) tcount[s] = 0;
    __syncwarp();
    const unsigned lt = lanemask_lt();
    int cnt = 0, q0 = 0;
    bool fail = false;
    int maxa = 0;
    const int pb_begin = Bp[j], pb_end = Bp[j + 1];
    for (int pb = pb_begin; pb < pb_end && !fail; pb++) {
        const int k = Bi[pb];
        const int ab = Ap[k], ae = Ap[k + 1];
        maxa = max(maxa, ae - ab);
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
            if (active) {
                tpl_pos[(size_t)c * TPL_UB + q0 + (pa - ab)] = (unsigned char)posof[slot];
                if (posof[slot] < TPL_CAP) tcount[posof[slot]]++;      // distinct rows in a step: no two lanes share a counter
            }
            cnt += __popc(newmask);
            if (cnt > TPL_CAP) { fail = true; break; }
            __syncwarp();
        }
        q0 += ae - ab;
    }
    if (lane == 0) tpl_cnt[c] = fail ? -1 : cnt;
    if (fail || !tpl_soa) return;
    // ---- inverted form for k_num_soa -----------------------------------------------------------
    // Rows are taken in chunks of 32 (they leave k_num_soa as 256-byte runs); inside a chunk every row
    // is given to one of SOA_WARPS warps, longest list first to the least loaded warp, and the term
    // lists of a warp's rows are laid end to end: the warp then walks ONE flat list per chunk, eight
    // terms at a time, whatever the rows' individual lengths.  tend[q] = row inside the chunk if term q
    // closes a row (its sum is complete), 0 otherwise.  tcount[r] holds the length of row r's list here.
    if (q0 > SOA_UB || cnt > SOA_CNT || maxa > soa_len_a || pb_end - pb_begin > soa_len_b) return;
    __shared__ int rstart[SOA_CNT];                  // where the list of row r starts in the reordered table
    __shared__ int rlen[SOA_CNT];
    for (int q = lane; q < SOA_TERMS; q += 32) {     // padding terms read a(0) * b(0) of the column; their sum is dropped
        tpl_terms[(size_t)c * SOA_TERMS + q] = 0;
        tpl_tend[(size_t)c * SOA_TERMS + q] = 0;
    }
    for (int r = lane; r < cnt; r += 32) rlen[r] = tcount[r];
    __syncwarp();
    const int nchunks = (cnt + 31) >> 5;
    int *wptr = tpl_wptr + (size_t)c * (SOA_CHUNKS * SOA_WARPS + 1);
    __shared__ unsigned char owner[SOA_CNT];
    __shared__ int wstart[SOA_CHUNKS * SOA_WARPS + 1];
    if (lane < nchunks) {                            // one lane per chunk: longest remaining row -> least loaded warp
        const int t0 = lane * 32, nrow = min(32, cnt - t0);
        int load[SOA_WARPS];
        for (int w = 0; w < SOA_WARPS; w++) load[w] = 0;
        unsigned taken = 0;
        for (int rank = 0; rank < nrow; rank++) {
            int best = -1, blen = -1;
            for (int r = 0; r < nrow; r++)
                if (!((taken >> r) & 1u) && rlen[t0 + r] > blen) { best = r; blen = rlen[t0 + r]; }
            taken |= 1u << best;
            int wmin = 0;
            for (int w = 1; w < SOA_WARPS; w++) if (load[w] < load[wmin]) wmin = w;
            owner[t0 + best] = (unsigned char)wmin;
            load[wmin] += blen;
        }
        for (int w = 0; w < SOA_WARPS; w++)
            wstart[lane * SOA_WARPS + w] = (load[w] + SOA_BATCH - 1) / SOA_BATCH * SOA_BATCH;   // whole steps only
    }
    __syncwarp();
    if (lane == 0) {                                 // where every (chunk, warp) list starts
        int run = 0;
        for (int k = 0; k < nchunks * SOA_WARPS; k++) { const int v = wstart[k]; wstart[k] = run; run += v; }
        for (int k = nchunks * SOA_WARPS; k <= SOA_CHUNKS * SOA_WARPS; k++) wstart[k] = run;
    }
    __syncwarp();
    for (int k = lane; k <= SOA_CHUNKS * SOA_WARPS; k += 32) wptr[k] = wstart[k];
    if (lane < nchunks) {
        const int t0 = lane * 32, nrow = min(32, cnt - t0);
        for (int w = 0; w < SOA_WARPS; w++) {
            int run = wstart[lane * SOA_WARPS + w];
            for (int r = 0; r < nrow; r++) if (owner[t0 + r] == w) { rstart[t0 + r] = run; run += rlen[t0 + r]; }
        }
    }
    __syncwarp();
    for (int r = lane; r < cnt; r += 32) tcount[r] = rstart[r];      // from here on: the fill cursor of row r
    __syncwarp();
    q0 = 0;
    bool far = false;
    for (int pb = pb_begin; pb < pb_end; pb++) {     // the same walk again: products in the reference's order
        const int k = Bi[pb];
        const int ab = Ap[k], ae = Ap[k + 1];
        const int d = pb - pb_begin;
        for (int pa0 = ab; pa0 < ae; pa0 += 32) {
            const int pa = pa0 + lane;
            if (pa < ae) {
                const int e = pa - ab;
                const int r = tpl_pos[(size_t)c * TPL_UB + q0 + e];
                const int idx = tcount[r]++;         // distinct rows in a step
                tpl_terms[(size_t)c * SOA_TERMS + idx] = soa_term(e, d, k - j);
                if (k - j >= SOA_DK_LIM || k - j < -SOA_DK_LIM) far = true;
                if (idx == rstart[r] + rlen[r] - 1)                       // the row's last term: 1 + row inside the chunk
                    tpl_tend[(size_t)c * SOA_TERMS + idx] = (unsigned char)((r & 31) + 1);
            }
            __syncwarp();
        }
        q0 += ae - ab;
    }
    if (!__any_sync(0xffffffffu, far) && lane == 0) tpl_soa[c] = 1;      // a column of A too far from j for the 20-bit field: k_num_tpl
}

// ---- verification and hand-out, one thread per column ----------------------------------------------
// Entry-by-entry comparison of the column with its class representative (a hash collision only sends
// the column to the general kernels).  Columns of a class with a template: cnt[j] is known and the
// general symbolic phase skips them (ub 0).  The warp (32 consecutive columns = one block of
// k_num_soa) also decides who forms them: a class with at least SOA_MIN_LANES columns in the block and
// the tables for it -> k_num_soa (mode 1); the rest (grid boundaries, long columns) go to the list
// k_num_tpl walks, one warp per column.  cb[j] <- dense class id, or -1.
__global__ void __launch_bounds__(256)
k_cls_verify_apply(int n, const csi *__restrict__ Bp, const csi *__restrict__ Bi, const int *__restrict__ ca,
                   ClsTable t, const int *__restrict__ tpl_cnt, int *__restrict__ cb, int *__restrict__ cnt,
                   int *__restrict__ ub, const unsigned char *__restrict__ tpl_soa, unsigned char *__restrict__ mode,
                   int *__restrict__ left_list)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int c = -1;
    if (j < n && !t.info[1]) {                       // the flag is final by now: a cached load
        const int slot = cb[j];
        bool ok = slot >= 0;
        if (ok) {
            const int r = t.rep[slot];
            if (r != j) {
                const int b = Bp[j], len = Bp[j + 1] - b, br = Bp[r];
                ok = (Bp[r + 1] - br) == len;
                // four entries per round, their loads in flight together (one entry per round with an early
                // exit made this a chain of dependent loads: 82 % of its samples were long-scoreboard stalls)
                for (int e = 0; ok && e < len; e += 4) {
                    int k[4], kr[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int eu = min(e + u, len - 1);            // past the end: the last entry again
                        k[u] = Bi[b + eu];
                        kr[u] = Bi[br + eu];
                    }
                    int ck[4], ckr[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) { ck[u] = ca[k[u]]; ckr[u] = ca[kr[u]]; }
#pragma unroll
                    for (int u = 0; u < 4; u++) ok &= (k[u] - j == kr[u] - r) && ck[u] == ckr[u];
                }
            }
            if (ok) c = t.dense[slot];
        }
    }
    const int tc = c >= 0 ? tpl_cnt[c] : -1;
    const bool hit = tc >= 0;
    if (!hit) c = -1;
    if (j < n) {
        cb[j] = c;
        if (hit) { cnt[j] = tc; ub[j] = 0; }
    }
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0 && m) atomicAdd(t.info + 3, __popc(m));
    bool soa = false;
    if (tpl_soa) {
        const unsigned same = __match_any_sync(0xffffffffu, c);
        soa = hit && tpl_soa[c] && __popc(same) >= SOA_MIN_LANES;
    }
    const unsigned left = __ballot_sync(0xffffffffu, hit && !soa);
    int base = 0;
    if (lane == 0 && left) base = atomicAdd(t.info + 4, __popc(left));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (hit && !soa) left_list[base + __popc(left & lanemask_lt())] = j;
    if (j < n) mode[j] = soa ? 1 : 0;
}

// ---- numeric on templates, one warp per column ----------------------------------------------------
// Shared memory per warp: TPL_CAP accumulators + the staged (first, length, B value) of 32 columns of A.
// The kernel is latency-bound unless loads are kept in flight: the values of TPL_BATCH columns of A
// (and their positions) are fetched together before the read-modify-write steps that consume them,
// and the B entries of the warp's NEXT column are fetched while the current one is accumulated.
constexpr int TPL_PER_WARP = TPL_CAP * 8 + 32 * 16;
template <bool VALUES, int TPL_BATCH, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_num_tpl(const int *__restrict__ left_list, const int *__restrict__ left_count, const int *__restrict__ cb,
          const int *__restrict__ tpl_cnt,
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

    // the columns of the list (every one has a template: cb[j] >= 0), strided over the warps
    const int n = *left_count;
    int ix = blockIdx.x * 8 + wid;
    int j = ix < n ? left_list[ix] : 0;
    int c = -1, pb_begin = 0, pb_end = 0;
    int4 st = make_int4(0, 0, 0, 0);
    if (ix < n) {
        c = cb[j];
        if (c >= 0) { pb_begin = Bp[j]; pb_end = Bp[j + 1]; if (VALUES) st = fetch(pb_begin + lane, pb_end); }
    }
    while (ix < n) {
        // the next column of this warp: its class and first 32 B entries travel while this one is summed
        const int ixn = ix + nwarps;
        const int jn = ixn < n ? left_list[ixn] : 0;
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
                    if (pb0 == pb_begin && ixn < n) {
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
            } else if (ixn < n) {
                cn = cb[jn];
            }
            const int *rows = tpl_rows + (size_t)c * TPL_CAP;
            for (int t = lane; t < cnt; t += 32) {
                Ci[out + t] = j + rows[t];
                if (VALUES) Cx[out + t] = vals[t];
            }
            __syncwarp();
        } else if (ixn < n) {
            cn = cb[jn];
            if (VALUES && cn >= 0) { pbn_begin = Bp[jn]; pbn_end = Bp[jn + 1]; stn = fetch(pbn_begin + lane, pbn_end); }
        }
        ix = ixn; j = jn; c = cn; pb_begin = pbn_begin; pb_end = pbn_end; st = stn;
    }
}

// ---- entry-major copies ------------------------------------------------------------------------------
// xT[e * n + k] = the e-th stored value of column k (columns shorter than `len` leave their slots
// unset; they are never read).  One thread per column: the strided reads of consecutive e share
// sectors through L1, the writes are coalesced.  Built once per handle.
__global__ void k_soa_build(int n, int len, const csi *__restrict__ Ap, const double *__restrict__ Ax,
                            double *__restrict__ xT)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int b = Ap[k], l = min(len, Ap[k + 1] - b);
    for (int e = 0; e < l; e++) xT[(size_t)e * n + k] = Ax[b + e];
}

__global__ void k_max_col_len(int n, const csi *__restrict__ Ap, int *out)
{
    int mx = 0;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) mx = max(mx, Ap[k + 1] - Ap[k]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, mx);
}

// ---- numeric on templates, one LANE per column ------------------------------------------------------
// Columns of one class do the same thing to different data, so 32 of them run in lock step: lane L
// owns column j = 32 g + L and walks the class's term lists -- for every row t of the column the
// products b * a that land on it, in the reference's order -- with both operands read from the
// entry-major copies, where the 32 lanes' operands are 32 consecutive doubles (one coalesced 256-byte
// request each).  The sum of a row lives in a register; nothing is accumulated in shared memory, no
// row index is read and the control flow is uniform.  The eight warps of a CTA share the 32 columns
// and split the rows, so they read the same lines of A and B through L1: the CTA's working set is
// the ~100 KB of A the 32 columns touch, which is why few CTAs are resident per SM (the grid is sized
// for it) and shared memory is kept small -- rows are staged 32 at a time and leave as 256-byte runs.
// Columns whose class has fewer than SOA_MIN_LANES members inside the 32-column block (grid
// boundaries) are left to k_num_tpl (k_tpl_apply makes that choice and lists them).
struct __align__(16) SoaTables {
    int terms[SOA_TERMS];
    unsigned char tend[SOA_TERMS];
    int wptr[SOA_CHUNKS * SOA_WARPS + 1];
    int rows[SOA_CNT];
    int out[32], mcol[32], mlane[32];                       // the member columns of the pass: Cp, column, lane
};
constexpr int SOA_SMEM = (int)sizeof(SoaTables) + 2 * 32 * 33 * 8;

template <int MINB>
__global__ void __launch_bounds__(SOA_THREADS, MINB)
k_num_soa(int n, const int *__restrict__ cb, const int *__restrict__ tpl_cnt, const unsigned char *__restrict__ mode,
          const int *__restrict__ tpl_terms, const unsigned char *__restrict__ tpl_tend, const int *__restrict__ tpl_wptr,
          const int *__restrict__ tpl_rows, const double *__restrict__ AxT, const double *__restrict__ BxT, int nA, int nB,
          const csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx, int stride, int *__restrict__ tickets)
{
    constexpr int NW = SOA_WARPS, BATCH = SOA_BATCH;
    static_assert(BATCH == 8 && SOA_TERMS % 8 == 0, "one 8-byte word of row marks per step");
    extern __shared__ __align__(16) unsigned char sm_raw[];
    SoaTables &tb = *reinterpret_cast<SoaTables *>(sm_raw);
    double *sCbuf = reinterpret_cast<double *>(sm_raw + sizeof(SoaTables));    // 2 x [lane][33]: 32 rows of 32 columns
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    int cur = -1;                                          // class whose tables are loaded (uniform over the CTA)
    const int nblocks = (n + 31) >> 5;
    // Order of the 32-column blocks.  stride == 0: round robin over the CTAs (A/B switch).  Otherwise the CTAs
    // resident on one SM work side by side: super-group G = 4 slots of `stride` consecutive blocks, handed
    // out by tickets[G]; SM s takes the super-groups s, s + #SMs, ... so the whole grid still sweeps the
    // matrix as one moving window (L2), while neighbouring column blocks -- which read neighbouring lines
    // of A -- share one SM's L1: 2.77 -> 2.49 ms on the 27-point stencil (the kernel waits on the A loads
    // that miss L1).  A contiguous range of blocks per SM instead loses the common window: 2.77 vs 2.92.
    // Nothing depends on where a CTA really runs: super-groups nobody owns (missing SM ids, an SM with
    // fewer CTAs) are found by the scan at the end, which also balances the tail.
    __shared__ int s_g;
    constexpr int NSLOT = 4;
    const int per_sg = stride * NSLOT, nG = stride > 0 ? (nblocks + per_sg - 1) / per_sg : 0;
    unsigned smid = 0, nsmid = 1;
    asm("mov.u32 %0, %%smid;" : "=r"(smid));
    asm("mov.u32 %0, %%nsmid;" : "=r"(nsmid));
    int G = (int)smid, slot = -1, pstep = 0, rr = blockIdx.x;               // thread 0's cursor
    int scanG = 0;                                                           // CTA-uniform: where the final scan stands
    bool own_done = false;
    for (;;) {
        int g;
        if (stride == 0) {
            g = rr; rr += gridDim.x;
            if (g >= nblocks) break;
        } else {
            __syncthreads();
            if (tid == 0) {
                int gg = -1;
                while (gg < 0) {
                    if (slot >= 0) {
                        if (pstep < stride) { const int t = G * per_sg + slot * stride + pstep++; if (t < nblocks) { gg = t; break; } }
                        slot = -1;
                    }
                    if (own_done || G >= nG) break;
                    const int t = atomicAdd(&tickets[G], 1);
                    if (t < NSLOT) { slot = t; pstep = 0; } else G += (int)nsmid;
                }
                s_g = gg;
            }
            __syncthreads();
            g = s_g;
            while (g < 0 && scanG < nG) {                                    // the own range is done: anything left anywhere?
                own_done = true;
                const int t = scanG + tid;
                const int seen = t < nG ? *reinterpret_cast<volatile int *>(tickets + t) : NSLOT;
                if (!__syncthreads_or(seen < NSLOT)) { scanG += SOA_THREADS; continue; }
                if (tid == 0) {
                    s_g = -1;
                    for (int G2 = scanG; G2 < min(nG, scanG + SOA_THREADS); G2++) {
                        if (*reinterpret_cast<volatile int *>(tickets + G2) >= NSLOT) continue;
                        const int tk = atomicAdd(&tickets[G2], 1);
                        if (tk < NSLOT) { G = G2; slot = tk; pstep = 1; s_g = G2 * per_sg + tk * stride; break; }
                    }
                }
                __syncthreads();
                g = s_g;
                if (g < 0) scanG += SOA_THREADS;                             // others took them meanwhile
                else if (g >= nblocks) g = -1;                               // a slot past the end: look again
            }
            if (g < 0) break;
        }
        const int j = g * 32 + lane;
        int c = (j < n && mode[j]) ? cb[j] : -1;           // k_tpl_apply decided which columns are formed here
        unsigned todo = __ballot_sync(0xffffffffu, c >= 0);
        while (todo) {                                     // one pass per class present in the block
            const int cc = __shfl_sync(0xffffffffu, c, __ffs(todo) - 1);
            const unsigned members = __ballot_sync(0xffffffffu, c == cc);
            todo &= ~members;
            const int nmem = __popc(members);
            const int cnt = tpl_cnt[cc];
            __syncthreads();                               // the previous pass has left the tables and sC
            if (cc != cur) {
                const int *wp = tpl_wptr + (size_t)cc * (SOA_CHUNKS * SOA_WARPS + 1);
                const int nterms = wp[SOA_CHUNKS * SOA_WARPS];
                for (int q = tid; q < nterms; q += SOA_THREADS) {
                    tb.terms[q] = tpl_terms[(size_t)cc * SOA_TERMS + q];
                    tb.tend[q] = tpl_tend[(size_t)cc * SOA_TERMS + q];
                }
                for (int r = tid; r <= SOA_CHUNKS * SOA_WARPS; r += SOA_THREADS) tb.wptr[r] = wp[r];
                for (int r = tid; r < cnt; r += SOA_THREADS) tb.rows[r] = tpl_rows[(size_t)cc * TPL_CAP + r];
                cur = cc;
            }
            const bool active = (members >> lane) & 1u;
            if (w == 0 && active) {                        // the member columns, densely: column and where it starts in C
                const int rank = __popc(members & lanemask_lt());
                tb.mcol[rank] = j;
                tb.out[rank] = Cp[j];
                tb.mlane[rank] = lane;
            }
            __syncthreads();
            const double *Aj = AxT + j, *Bj = BxT + j;     // this lane's column in the entry-major copies
            // opaque to the compiler from here on: an operand address is then ONE IMAD.WIDE (offset * 8 + pointer)
            // instead of the four-instruction 64-bit sum (base + (j + offset) * 8) it otherwise rebuilds per load --
            // 64 of the 167 instructions of a step of eight terms (profiles/r2_notes.md)
            asm volatile("" : "+l"(Aj), "+l"(Bj));
            int buf = 0;
            for (int t0 = 0; t0 < cnt; t0 += 32, buf ^= 1) {
                double *sC = sCbuf + buf * (32 * 33) + lane * 33;
                // shared-space byte address of sC[-1]: a completed row e (1-based) is stored at sCm1 + 8 e
                const unsigned sCm1 = (unsigned)__cvta_generic_to_shared(sC) - 8u;
                // this warp's share of the chunk: one flat list of terms, eight per step -- sixteen
                // independent loads in flight, then the sums in list order.  -0.0 is the exact additive
                // identity: the first product of a row lands as the reference's first-touch assignment
                // (csparse.py:1986), the rest are added in its order (:1988)
                const int Q1 = tb.wptr[(t0 >> 5) * NW + w + 1];
                if (active) {
                    double acc = -0.0;
                    for (int q = tb.wptr[(t0 >> 5) * NW + w]; q < Q1; q += BATCH) {
                        int tm[BATCH];
                        double a[BATCH], b[BATCH];
#pragma unroll
                        for (int u = 0; u < BATCH; u += 4) {
                            const int4 four = *reinterpret_cast<const int4 *>(&tb.terms[q + u]);
                            tm[u] = four.x; tm[u + 1] = four.y; tm[u + 2] = four.z; tm[u + 3] = four.w;
                        }
                        const unsigned long long marks = *reinterpret_cast<const unsigned long long *>(&tb.tend[q]);
#pragma unroll
                        for (int u = 0; u < BATCH; u++) {
                            b[u] = __ldg(Bj + ((tm[u] >> 6) & 63) * nB);
                            a[u] = __ldg(Aj + ((tm[u] & 63) * nA + (tm[u] >> 12)));
                        }
#pragma unroll
                        for (int u = 0; u < BATCH; u++) {
                            acc = __dadd_rn(acc, __dmul_rn(b[u], a[u]));
                            const int e = (int)((marks >> (8 * u)) & 0xffull);
                            if (e) {                                        // warp-uniform: the row is complete
                                asm volatile("st.shared.f64 [%0], %1;" :: "r"(sCm1 + 8u * (unsigned)e), "d"(acc) : "memory");
                                acc = -0.0;
                            }
                        }
                    }
                }
                __syncthreads();
                // every warp writes 32 consecutive rows of one member column: a 256-byte run of Cx.  The
                // next chunk is summed into the other buffer, so one barrier per chunk is enough.
                const int t = t0 + lane;
                if (t < cnt) {
                    const int rt = tb.rows[t];
                    const double *src = sCbuf + buf * (32 * 33) + lane;
                    for (int mi = w; mi < nmem; mi += NW) {
                        const int o = tb.out[mi] + t;
                        Cx[o] = src[tb.mlane[mi] * 33];
                        Ci[o] = tb.mcol[mi] + rt;
                    }
                }
            }
        }
    }
}

}  // namespace csb
