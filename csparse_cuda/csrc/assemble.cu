// assemble.cu -- the callers and data formats either side of the hot path:
//
//   cs_add      csparse.py:163-192    C = alpha*A + beta*B
//   cs_norm     csparse.py:1647-1663  1-norm (largest column sum of |x|)
//   cs_compress csparse.py:647-672    triplet -> compressed column
//   cs_dupl     csparse.py:1035-1063  sum duplicate entries
//   cs_fkeep    csparse.py:1172-1196  with the fixed predicates of cs_dropzeros (:1024),
//                                     cs_droptol (:1007) and "off-diagonal" (csparse_test.py Dropdiag)
//   cs_permute  csparse.py:1666-1693  C = P A Q
//   cs_symperm  csparse.py:2220-2255  C = P A P' (upper triangular part)
//
// They are built from the hot path's own pieces:
//   * cs_add and cs_dupl ARE cs_scatter loops, so they run on the SpGEMM kernels:
//     alpha*A + beta*B = [A B] * [alpha I; beta I] and dupl(A) = A * I.  Column j of the product
//     scatters A(:,j) with alpha and then B(:,j) with beta -- the reference's own sequence --
//     so pattern order and rounding are the reference's.
//   * cs_compress and cs_symperm are stable counting sorts by column: radix.cu.
//   * cs_fkeep / cs_permute are a flag or length pass, the look-back scan, and one copy.
#include "common.cuh"
#include <string.h>

namespace csb {

int multiply_impl(csb200_mat *A, csb200_mat *B, csb200_mat **out, bool ordered);
int mat_is_canonical(csb200_mat *A, int *out);
int mat_alloc(csi m, csi n, long long nnz, bool has_x, csb200_mat **out);

// ---- small builders ------------------------------------------------------------------
// M = [A B]: p = [Ap, nnzA + Bp[1..]]
__global__ void k_hcat_p(int nA, int nB, const csi *__restrict__ Ap, const csi *__restrict__ Bp, int nnzA,
                         csi *__restrict__ Mp)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= nA) Mp[j] = Ap[j];
    if (j >= 1 && j <= nB) Mp[nA + j] = nnzA + Bp[j];
}

// S = [alpha I; beta I] (2n x n): column j holds (j, alpha), (n + j, beta)
__global__ void k_add_rhs(int n, double alpha, double beta, csi *__restrict__ p, csi *__restrict__ i, double *x)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) {
        p[j] = 2 * j;
        i[2 * j] = j;
        i[2 * j + 1] = n + j;
        if (x) { x[2 * j] = alpha; x[2 * j + 1] = beta; }
    }
    if (j == n) p[n] = 2 * n;
}

__global__ void k_identity(int n, csi *__restrict__ p, csi *__restrict__ i, double *__restrict__ x)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) { p[j] = j; i[j] = j; x[j] = 1.0; }
    if (j == n) p[n] = n;
}

// ---- cs_add on canonical operands ---------------------------------------------------------
// When every column of A and of B is strictly increasing (no duplicates), column j of
// alpha*A + beta*B is A(:,j) in order followed by the rows of B(:,j) that A(:,j) lacks, in
// order (that is what the two cs_scatter calls of csparse.py:186-187 discover), with values
// alpha*a, alpha*a + beta*b, beta*b -- each product rounded, then the sum.  Membership is a
// binary search in the sorted column of A.  G = 1: one thread per column (short columns);
// G = 32: one warp per column.
template <int G>
__device__ __forceinline__ int add_find(const csi *__restrict__ Ai, int lo, int hi, int r)
{
    while (lo < hi) {                        // first position in [lo, hi) with Ai[pos] >= r
        const int mid = (lo + hi) >> 1;
        if (Ai[mid] < r) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int G>
__global__ void __launch_bounds__(256)
k_add_count(int n, const csi *__restrict__ Ap, const csi *__restrict__ Ai,
            const csi *__restrict__ Bp, const csi *__restrict__ Bi, int *__restrict__ cnt)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = t / G, lane = t % G;
    if (j >= n) return;
    const int a0 = Ap[j], a1 = Ap[j + 1], b0 = Bp[j], b1 = Bp[j + 1];
    int fresh = 0;
    for (int q = b0 + lane; q < b1; q += G) {
        const int r = Bi[q];
        const int pos = add_find<G>(Ai, a0, a1, r);
        fresh += !(pos < a1 && Ai[pos] == r);
    }
    if (G == 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fresh += __shfl_xor_sync(0xffffffffu, fresh, o);
    }
    if (lane == 0) cnt[j] = (a1 - a0) + fresh;
}

template <int G, bool VALUES>
__global__ void __launch_bounds__(256)
k_add_fill(int n, double alpha, double beta,
           const csi *__restrict__ Ap, const csi *__restrict__ Ai, const double *__restrict__ Ax,
           const csi *__restrict__ Bp, const csi *__restrict__ Bi, const double *__restrict__ Bx,
           const csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = t / G, lane = t % G;
    if (j >= n) return;
    const int a0 = Ap[j], a1 = Ap[j + 1], b0 = Bp[j], b1 = Bp[j + 1];
    const int c0 = Cp[j];
    for (int q = a0 + lane; q < a1; q += G) {
        Ci[c0 + (q - a0)] = Ai[q];
        if (VALUES) Cx[c0 + (q - a0)] = __dmul_rn(alpha, Ax[q]);
    }
    if (G == 32) __syncwarp();
    int out = c0 + (a1 - a0);                                   // next free slot (uniform in a warp)
    for (int q0 = b0; q0 < b1; q0 += G) {
        const int q = q0 + lane;
        const bool valid = q < b1;
        int r = 0, pos = a1;
        bool found = false;
        if (valid) {
            r = Bi[q];
            pos = add_find<G>(Ai, a0, a1, r);
            found = pos < a1 && Ai[pos] == r;
        }
        const double bv = (VALUES && valid) ? __dmul_rn(beta, Bx[q]) : 0.0;
        if (G == 32) {
            const unsigned fresh = __ballot_sync(0xffffffffu, valid && !found);
            if (valid) {
                if (found) { if (VALUES) Cx[c0 + (pos - a0)] = __dadd_rn(Cx[c0 + (pos - a0)], bv); }
                else {
                    const int k = out + __popc(fresh & lanemask_lt());
                    Ci[k] = r;
                    if (VALUES) Cx[k] = bv;
                }
            }
            out += __popc(fresh);
        } else if (valid) {
            if (found) { if (VALUES) Cx[c0 + (pos - a0)] = __dadd_rn(Cx[c0 + (pos - a0)], bv); }
            else { Ci[out] = r; if (VALUES) Cx[out] = bv; out++; }
        }
    }
}

// ---- cs_norm ---------------------------------------------------------------------------
// One thread per column, entries added in storage order (the reference's rounding); the
// maximum is kept as the bit pattern of a non-negative double.  A NaN column sum never
// wins, as in the reference's max(norm, s).
__global__ void k_norm1(int n, const csi *__restrict__ Ap, const double *__restrict__ Ax,
                        unsigned long long *best)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    double s = 0.0;
    if (j < n)
        for (int p = Ap[j]; p < Ap[j + 1]; p++) s = __dadd_rn(s, fabs(Ax[p]));
    double w = s > 0.0 ? s : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, w, o); if (t > w) w = t; }
    if ((threadIdx.x & 31) == 0 && w > 0.0) atomicMax(best, (unsigned long long)__double_as_longlong(w));
}

// ---- cs_fkeep / cs_symperm flags ----------------------------------------------------------
enum { KEEP_NONZERO = 0, KEEP_TOL = 1, KEEP_OFFDIAG = 2, KEEP_UPPER = 3, KEEP_SHORTCOL = 4 };

__global__ void k_keep_flags(int n, long long nnz, const csi *__restrict__ Ap, const csi *__restrict__ Ai,
                             const double *__restrict__ Ax, int mode, double tol, int *__restrict__ flag)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    int keep;
    if (mode == KEEP_NONZERO || mode == KEEP_TOL) {
        const double a = Ax ? Ax[p] : 1.0;                       // pattern-only: aij taken as 1 (:1188)
        keep = mode == KEEP_NONZERO ? (a != 0.0) : (fabs(a) > tol);
    } else {
        const int j = upper_row(Ap, 0, n, (int)p);               // column holding entry p
        if (mode == KEEP_SHORTCOL) keep = (double)(Ap[j + 1] - Ap[j]) <= tol;   // cs_amd's dense-column drop (:236-249)
        else keep = mode == KEEP_OFFDIAG ? (Ai[p] != j) : (Ai[p] <= j);
    }
    flag[p] = keep;
}

// pos = exclusive scan of the flags (nnz + 1 entries); entry p is kept iff pos[p+1] > pos[p]
__global__ void k_compact(int n, long long nnz, const csi *__restrict__ Ap, const csi *__restrict__ Ai,
                          const double *__restrict__ Ax, const int *__restrict__ pos,
                          csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p <= n) Cp[p] = pos[Ap[p]];
    if (p >= nnz) return;
    const int q = pos[p];
    if (pos[p + 1] > q) {
        Ci[q] = Ai[p];
        if (Cx) Cx[q] = Ax[p];
    }
}

// upper-triangular entries of A, relabelled: key = max(i2, j2) (the output column), a = min
__global__ void k_symperm_keys(int n, long long nnz, const csi *__restrict__ Ap, const csi *__restrict__ Ai,
                               const double *__restrict__ Ax, const csi *__restrict__ pinv,
                               const int *__restrict__ pos, int *__restrict__ key, int *__restrict__ a,
                               double *__restrict__ v)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nnz) return;
    const int q = pos[p];
    if (pos[p + 1] > q) {
        const int j = upper_row(Ap, 0, n, (int)p);
        const int i = Ai[p];
        const int i2 = pinv ? pinv[i] : i, j2 = pinv ? pinv[j] : j;
        key[q] = max(i2, j2);
        a[q] = min(i2, j2);
        if (v) v[q] = Ax[p];
    }
}

// ---- cs_permute ------------------------------------------------------------------------------
__global__ void k_perm_lens(int n, const csi *__restrict__ Ap, const csi *__restrict__ q, int *__restrict__ len)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) { const int j = q ? q[k] : k; len[k] = Ap[j + 1] - Ap[j]; }
}

// G = 1: one thread per output column (short columns), G = 32: one warp per output column
template <int G>
__global__ void __launch_bounds__(256)
k_perm_copy(int n, const csi *__restrict__ Ap, const csi *__restrict__ Ai,
            const double *__restrict__ Ax, const csi *__restrict__ pinv, const csi *__restrict__ q,
            const csi *__restrict__ Cp, csi *__restrict__ Ci, double *__restrict__ Cx)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int k = (int)(t / G), lane = (int)(t % G);
    if (k >= n) return;
    const int j = q ? q[k] : k;
    const int s0 = Ap[j], len = Ap[j + 1] - s0, d0 = Cp[k];
    for (int e = lane; e < len; e += G) {
        const int i = Ai[s0 + e];
        Ci[d0 + e] = pinv ? pinv[i] : i;
        if (Cx) Cx[d0 + e] = Ax[s0 + e];
    }
}

__global__ void k_check_range(const csi *__restrict__ v, long long count, int bound, int *bad)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count && (v[k] < 0 || v[k] >= bound)) *bad = 1;
}

static int check_range(const csi *d_v, long long count, int bound, const char *what)
{
    if (count == 0) return CSB200_OK;
    DevBuf<int> bad;
    CSB_TRY(bad.alloc(1));
    CSB_CUDA(cudaMemsetAsync(bad.ptr, 0, sizeof(int), stream()));
    k_check_range<<<ceil_div(count, 256), 256, 0, stream()>>>(d_v, count, bound, bad.ptr);
    CSB_LAUNCHED();
    int h = 0;
    CSB_CUDA(cudaMemcpyAsync(&h, bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, stream()));
    CSB_CUDA(cudaStreamSynchronize(stream()));
    if (h) return set_error(CSB200_ERR_INDEX, "%s: index outside [0, %d)", what, bound);
    return CSB200_OK;
}

// host permutation (or NULL) -> device copy, range-checked
static int upload_perm(const csi *h, int count, int bound, DevBuf<csi> &d, const char *what)
{
    if (!h || count == 0) return CSB200_OK;
    CSB_TRY(d.alloc((size_t)count));
    CSB_CUDA(cudaMemcpyAsync(d.ptr, h, (size_t)count * sizeof(csi), cudaMemcpyHostToDevice, stream()));
    return check_range(d.ptr, count, bound, what);
}

struct MatGuard {                    // frees a temporary handle on every exit path
    csb200_mat *m = nullptr;
    ~MatGuard() { if (m) csb200_mat_free(m); }
};

static int empty_like(csi m, csi n, bool has_x, csb200_mat **out)
{
    csb200_mat *C = nullptr;
    CSB_TRY(mat_alloc(m, n, 0, has_x, &C));
    cudaError_t e = cudaMemsetAsync(C->p, 0, ((size_t)n + 1) * sizeof(csi), stream());
    if (e != cudaSuccess) { csb200_mat_free(C); return set_error(CSB200_ERR_CUDA, "memset: %s", cudaGetErrorString(e)); }
    *out = C;
    return CSB200_OK;
}

}  // namespace csb

using namespace csb;

extern "C" {

// ---- cs_add ------------------------------------------------------------------------------------
int csb200_add(csb200_mat *A, csb200_mat *B, double alpha, double beta, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!A || !B || !C) return set_error(CSB200_ERR_ARG, "cs_add: null argument");
    *C = nullptr;
    if (A->m != B->m || A->n != B->n) return set_error(CSB200_ERR_ARG, "cs_add: dimension mismatch");
    const csi m = A->m, n = A->n;
    if ((long long)A->nnz + B->nnz > 0x7fffffffLL || (long long)n * 2 > 0x7fffffffLL)
        return set_error(CSB200_ERR_OVERFLOW, "cs_add: nnz(A) + nnz(B) does not fit int32");
    const bool values = A->x && B->x;                                    // csparse.py:180
    cudaStream_t s = stream();
    int ca = 0, cb = 0;
    CSB_TRY(mat_is_canonical(A, &ca));
    CSB_TRY(mat_is_canonical(B, &cb));
    if (ca && cb && !tls().add_force_spgemm) {
        // both operands sorted without duplicates: membership by binary search, two passes
        DevBuf<int> cnt;
        DevBuf<long long> total;
        DevBuf<int> dmax;
        CSB_TRY(cnt.alloc((size_t)n + 1));
        CSB_TRY(total.alloc(1));
        MatGuard R;
        csb200_mat *Cm = new csb200_mat();
        R.m = Cm;
        Cm->m = m; Cm->n = n; Cm->device = A->device;
        CSB_TRY(dev_alloc(&Cm->p, (size_t)n + 1 + MAT_PAD));
        const long long avg = n > 0 ? (A->nnz + B->nnz) / n : 0;
        const bool warp = avg > 24;                                      // long columns: a warp each
        if (n > 0) {
            if (warp) k_add_count<32><<<ceil_div((long long)n * 32, 256), 256, 0, s>>>(n, A->p, A->i, B->p, B->i, cnt.ptr);
            else      k_add_count<1><<<ceil_div(n, 256), 256, 0, s>>>(n, A->p, A->i, B->p, B->i, cnt.ptr);
            CSB_LAUNCHED();
        }
        CSB_TRY(launch_excl_scan(Cm->p, cnt.ptr, n, total.ptr, nullptr));
        long long h_total = 0;
        CSB_CUDA(cudaMemcpyAsync(&h_total, total.ptr, sizeof(h_total), cudaMemcpyDeviceToHost, s));
        CSB_CUDA(cudaStreamSynchronize(s));
        Cm->nnz = h_total;
        const size_t cap = (size_t)(h_total > 0 ? h_total : 1) + MAT_PAD;
        CSB_TRY(dev_alloc(&Cm->i, cap));
        if (values) CSB_TRY(dev_alloc(&Cm->x, cap));
        if (h_total > 0) {
            const int grid = warp ? ceil_div((long long)n * 32, 256) : ceil_div(n, 256);
#define ADD_FILL(G, V) k_add_fill<G, V><<<grid, 256, 0, s>>>(n, alpha, beta, A->p, A->i, A->x, B->p, B->i, B->x, \
                                                            Cm->p, Cm->i, Cm->x)
            if (warp) { if (values) ADD_FILL(32, true); else ADD_FILL(32, false); }
            else      { if (values) ADD_FILL(1, true); else ADD_FILL(1, false); }
#undef ADD_FILL
            CSB_LAUNCHED();
        }
        *C = R.m;
        R.m = nullptr;
        return CSB200_OK;
    }
    MatGuard M, S;
    CSB_TRY(mat_alloc(m, 2 * n, A->nnz + B->nnz, values, &M.m));
    CSB_TRY(mat_alloc(2 * n, n, 2LL * n, values, &S.m));
    k_hcat_p<<<ceil_div((long long)n + 1, 256), 256, 0, s>>>(n, n, A->p, B->p, (int)A->nnz, M.m->p);
    CSB_LAUNCHED();
    if (A->nnz) CSB_CUDA(cudaMemcpyAsync(M.m->i, A->i, (size_t)A->nnz * sizeof(csi), cudaMemcpyDeviceToDevice, s));
    if (B->nnz) CSB_CUDA(cudaMemcpyAsync(M.m->i + A->nnz, B->i, (size_t)B->nnz * sizeof(csi), cudaMemcpyDeviceToDevice, s));
    if (values) {
        if (A->nnz) CSB_CUDA(cudaMemcpyAsync(M.m->x, A->x, (size_t)A->nnz * sizeof(double), cudaMemcpyDeviceToDevice, s));
        if (B->nnz) CSB_CUDA(cudaMemcpyAsync(M.m->x + A->nnz, B->x, (size_t)B->nnz * sizeof(double), cudaMemcpyDeviceToDevice, s));
    }
    k_add_rhs<<<ceil_div((long long)n + 1, 256), 256, 0, s>>>(n, alpha, beta, S.m->p, S.m->i, S.m->x);
    CSB_LAUNCHED();
    return multiply_impl(M.m, S.m, C, true);
}

int csb200_add_force_path(int path)
{
    if (path < 0 || path > 1) return set_error(CSB200_ERR_ARG, "bad cs_add path");
    tls().add_force_spgemm = path;
    return CSB200_OK;
}

// ---- cs_norm -----------------------------------------------------------------------------------
int csb200_norm(const csb200_mat *A, double *norm)
{
    ArenaScope arena_scope;
    if (!A || !norm || !A->x) return set_error(CSB200_ERR_ARG, "cs_norm: null argument or no values");
    *norm = 0.0;
    if (A->n == 0) return CSB200_OK;
    DevBuf<unsigned long long> best;
    CSB_TRY(best.alloc(1));
    CSB_CUDA(cudaMemsetAsync(best.ptr, 0, sizeof(unsigned long long), stream()));
    k_norm1<<<ceil_div(A->n, 256), 256, 0, stream()>>>(A->n, A->p, A->x, best.ptr);
    CSB_LAUNCHED();
    unsigned long long h = 0;
    CSB_CUDA(cudaMemcpyAsync(&h, best.ptr, sizeof(h), cudaMemcpyDeviceToHost, stream()));
    CSB_CUDA(cudaStreamSynchronize(stream()));
    memcpy(norm, &h, sizeof(double));
    return CSB200_OK;
}

// ---- cs_compress -------------------------------------------------------------------------------
int csb200_compress_dev(csi m, csi n, csi nz, const csi *d_Ti, const csi *d_Tj, const double *d_Tx,
                        csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!C || m < 0 || n < 0 || nz < 0 || (nz > 0 && (!d_Ti || !d_Tj)))
        return set_error(CSB200_ERR_ARG, "cs_compress: bad arguments");
    *C = nullptr;
    if (nz == 0) return empty_like(m, n, d_Tx != nullptr, C);
    CSB_TRY(check_range(d_Ti, nz, m, "cs_compress: row index"));
    CSB_TRY(check_range(d_Tj, nz, n, "cs_compress: column index"));
    MatGuard R;
    CSB_TRY(mat_alloc(m, n, nz, d_Tx != nullptr, &R.m));
    CSB_TRY(stable_sort_by_key(nz, n, d_Tj, d_Ti, nullptr, 0, d_Tx, R.m->p, R.m->i, R.m->x));
    *C = R.m;
    R.m = nullptr;
    return CSB200_OK;
}

int csb200_compress(csi m, csi n, csi nz, const csi *Ti, const csi *Tj, const double *Tx, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!C || nz < 0 || (nz > 0 && (!Ti || !Tj))) return set_error(CSB200_ERR_ARG, "cs_compress: bad arguments");
    DevBuf<csi> dI, dJ;
    DevBuf<double> dX;
    if (nz > 0) {
        CSB_TRY(dI.alloc((size_t)nz));
        CSB_TRY(dJ.alloc((size_t)nz));
        CSB_CUDA(cudaMemcpyAsync(dI.ptr, Ti, (size_t)nz * sizeof(csi), cudaMemcpyHostToDevice, stream()));
        CSB_CUDA(cudaMemcpyAsync(dJ.ptr, Tj, (size_t)nz * sizeof(csi), cudaMemcpyHostToDevice, stream()));
        if (Tx) {
            CSB_TRY(dX.alloc((size_t)nz));
            CSB_CUDA(cudaMemcpyAsync(dX.ptr, Tx, (size_t)nz * sizeof(double), cudaMemcpyHostToDevice, stream()));
        }
    }
    int st = csb200_compress_dev(m, n, nz, dI.ptr, dJ.ptr, Tx ? (nz > 0 ? dX.ptr : (const double *)Tx) : nullptr, C);
    if (st == CSB200_OK) CSB_CUDA(cudaStreamSynchronize(stream()));      // the staging buffers die with this call
    return st;
}

// ---- cs_dupl -----------------------------------------------------------------------------------
int csb200_dupl(csb200_mat *A, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!A || !C || !A->x) return set_error(CSB200_ERR_ARG, "cs_dupl: null argument or no values");
    *C = nullptr;
    const csi n = A->n;
    int canon = 0;
    CSB_TRY(mat_is_canonical(A, &canon));
    if (canon) return csb200_mat_col_slice(A, 0, n, C);        // strictly increasing columns hold no duplicates
    MatGuard I;
    CSB_TRY(mat_alloc(n, n, n, true, &I.m));
    k_identity<<<ceil_div((long long)n + 1, 256), 256, 0, stream()>>>(n, I.m->p, I.m->i, I.m->x);
    CSB_LAUNCHED();
    return multiply_impl(A, I.m, C, true);
}

// ---- cs_fkeep with a fixed predicate -------------------------------------------------------------
int csb200_fkeep(const csb200_mat *A, int predicate, double tol, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!A || !C || predicate < KEEP_NONZERO || predicate > KEEP_SHORTCOL)
        return set_error(CSB200_ERR_ARG, "cs_fkeep: bad arguments");
    *C = nullptr;
    const long long nnz = A->nnz;
    if (nnz == 0) return empty_like(A->m, A->n, A->x != nullptr, C);
    cudaStream_t s = stream();
    DevBuf<int> flag, pos;
    DevBuf<long long> total;
    CSB_TRY(flag.alloc((size_t)nnz + 1));
    CSB_TRY(pos.alloc((size_t)nnz + 1));
    CSB_TRY(total.alloc(1));
    k_keep_flags<<<ceil_div(nnz, 256), 256, 0, s>>>(A->n, nnz, A->p, A->i, A->x, predicate, tol, flag.ptr);
    CSB_LAUNCHED();
    CSB_TRY(launch_excl_scan(pos.ptr, flag.ptr, (csi)nnz, total.ptr, nullptr));
    long long kept = 0;
    CSB_CUDA(cudaMemcpyAsync(&kept, total.ptr, sizeof(kept), cudaMemcpyDeviceToHost, s));
    CSB_CUDA(cudaStreamSynchronize(s));
    MatGuard R;
    CSB_TRY(mat_alloc(A->m, A->n, kept, A->x != nullptr, &R.m));
    k_compact<<<ceil_div(max(nnz, (long long)A->n + 1), 256), 256, 0, s>>>(A->n, nnz, A->p, A->i, A->x, pos.ptr,
                                                                          R.m->p, R.m->i, R.m->x);
    CSB_LAUNCHED();
    *C = R.m;
    R.m = nullptr;
    return CSB200_OK;
}

// ---- cs_permute ------------------------------------------------------------------------------------
int csb200_permute(const csb200_mat *A, const csi *pinv, const csi *q, int values, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!A || !C) return set_error(CSB200_ERR_ARG, "cs_permute: null argument");
    *C = nullptr;
    const csi m = A->m, n = A->n;
    const long long nnz = A->nnz;
    const bool has_x = values && A->x;
    if (nnz == 0 || n == 0) return empty_like(m, n, has_x, C);
    cudaStream_t s = stream();
    DevBuf<csi> d_pinv, d_q;
    CSB_TRY(upload_perm(pinv, m, m, d_pinv, "cs_permute: pinv"));
    CSB_TRY(upload_perm(q, n, n, d_q, "cs_permute: q"));
    DevBuf<int> len;
    DevBuf<long long> total;
    CSB_TRY(len.alloc((size_t)n + 1));
    CSB_TRY(total.alloc(1));
    MatGuard R;
    CSB_TRY(mat_alloc(m, n, nnz, has_x, &R.m));
    k_perm_lens<<<ceil_div(n, 256), 256, 0, s>>>(n, A->p, q ? d_q.ptr : nullptr, len.ptr);
    CSB_LAUNCHED();
    CSB_TRY(launch_excl_scan(R.m->p, len.ptr, n, total.ptr, nullptr));
    // C holds nnz(A) entries: a q that repeats columns (not a permutation) would make the copy
    // run past it -- the reference fails with IndexError there, this returns CSB200_ERR_INDEX
    long long h_total = 0;
    CSB_CUDA(cudaMemcpyAsync(&h_total, total.ptr, sizeof(h_total), cudaMemcpyDeviceToHost, s));
    CSB_CUDA(cudaStreamSynchronize(s));
    if (h_total > nnz)
        return set_error(CSB200_ERR_INDEX, "cs_permute: q is not a permutation (selected columns hold %lld entries, A has %lld)",
                         h_total, nnz);
    R.m->nnz = h_total;                                                   // fewer when q skips columns
    if (nnz / n > 12)
        k_perm_copy<32><<<ceil_div((long long)n * 32, 256), 256, 0, s>>>(n, A->p, A->i, has_x ? A->x : nullptr,
                                                                         pinv ? d_pinv.ptr : nullptr, q ? d_q.ptr : nullptr,
                                                                         R.m->p, R.m->i, R.m->x);
    else
        k_perm_copy<1><<<ceil_div(n, 256), 256, 0, s>>>(n, A->p, A->i, has_x ? A->x : nullptr,
                                                        pinv ? d_pinv.ptr : nullptr, q ? d_q.ptr : nullptr,
                                                        R.m->p, R.m->i, R.m->x);
    CSB_LAUNCHED();
    CSB_CUDA(cudaStreamSynchronize(s));                                   // d_pinv / d_q die with this call
    *C = R.m;
    R.m = nullptr;
    return CSB200_OK;
}

// ---- cs_symperm -------------------------------------------------------------------------------------
int csb200_symperm(const csb200_mat *A, const csi *pinv, int values, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!A || !C) return set_error(CSB200_ERR_ARG, "cs_symperm: null argument");
    *C = nullptr;
    const csi n = A->n;
    const long long nnz = A->nnz;
    const bool has_x = values && A->x;
    if (nnz == 0 || n == 0) return empty_like(n, n, has_x, C);
    cudaStream_t s = stream();
    DevBuf<csi> d_pinv;
    CSB_TRY(upload_perm(pinv, n, n, d_pinv, "cs_symperm: pinv"));
    DevBuf<int> flag, pos, key, a;
    DevBuf<double> v;
    DevBuf<long long> total;
    CSB_TRY(flag.alloc((size_t)nnz + 1));
    CSB_TRY(pos.alloc((size_t)nnz + 1));
    CSB_TRY(total.alloc(1));
    k_keep_flags<<<ceil_div(nnz, 256), 256, 0, s>>>(n, nnz, A->p, A->i, nullptr, KEEP_UPPER, 0.0, flag.ptr);
    CSB_LAUNCHED();
    CSB_TRY(launch_excl_scan(pos.ptr, flag.ptr, (csi)nnz, total.ptr, nullptr));
    long long kept = 0;
    CSB_CUDA(cudaMemcpyAsync(&kept, total.ptr, sizeof(kept), cudaMemcpyDeviceToHost, s));
    CSB_CUDA(cudaStreamSynchronize(s));
    if (kept == 0) return empty_like(n, n, has_x, C);
    CSB_TRY(key.alloc((size_t)kept));
    CSB_TRY(a.alloc((size_t)kept));
    if (has_x) CSB_TRY(v.alloc((size_t)kept));
    k_symperm_keys<<<ceil_div(nnz, 256), 256, 0, s>>>(n, nnz, A->p, A->i, has_x ? A->x : nullptr,
                                                      pinv ? d_pinv.ptr : nullptr, pos.ptr, key.ptr, a.ptr,
                                                      has_x ? v.ptr : nullptr);
    CSB_LAUNCHED();
    MatGuard R;
    CSB_TRY(mat_alloc(n, n, kept, has_x, &R.m));
    CSB_TRY(stable_sort_by_key(kept, n, key.ptr, a.ptr, nullptr, 0, has_x ? v.ptr : nullptr, R.m->p, R.m->i, R.m->x));
    CSB_CUDA(cudaStreamSynchronize(s));
    *C = R.m;
    R.m = nullptr;
    return CSB200_OK;
}

}  // extern "C"
