// api.cu -- the extern "C" surface of libcsparse_b200.so (include/csparse_b200.h):
// library state, matrix handles, and the host-buffer / device-buffer forms of
// cs_cumsum, cs_transpose, cs_gaxpy and cs_multiply.
#include "common.cuh"

#include <mutex>
#include <new>

namespace csb {

std::atomic<int64_t> g_launches{0};

ThreadState &tls()
{
    static thread_local ThreadState st;
    return st;
}

int set_error(int status, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls().err, sizeof(tls().err), fmt, ap);
    va_end(ap);
    return status;
}

int sm_count()
{
    static std::atomic<int> cached[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 148; }
    int v = cached[dev].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) { cudaGetLastError(); v = 148; }
        cached[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

Arena &arena()
{
    static thread_local Arena a;
    return a;
}

static int arena_grow(Arena &a, size_t bytes)
{
    // nothing of the block is handed out here (outermost level, offset 0); kernels of earlier
    // calls may still be reading it
    cudaStream_t s = stream();
    if (a.base) {
        cudaStreamSynchronize(a.last_stream);
        if (a.last_stream != s) cudaStreamSynchronize(s);
        cudaFree(a.base);
        cudaGetLastError();
        a.base = nullptr;
        a.cap = 0;
    }
    bytes += bytes / 8 + (1 << 20);
    bytes = (bytes + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {      // no room for a private block: the pool keeps serving
        cudaGetLastError();
        return CSB200_OK;
    }
    a.base = static_cast<char *>(p);
    a.cap = bytes;
    return CSB200_OK;
}

int arena_enter()
{
    Arena &a = arena();
    if (a.depth > 0) return CSB200_OK;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return CSB200_OK; }
    cudaStream_t s = stream();
    if (a.base && a.device != dev) {                  // the thread moved to another GPU: start over there
        int cur = dev;
        cudaSetDevice(a.device);
        cudaDeviceSynchronize();
        cudaFree(a.base);
        cudaGetLastError();
        cudaSetDevice(cur);
        a.base = nullptr; a.cap = 0; a.want = 0;
    }
    a.device = dev;
    if (a.base && a.last_stream != s) cudaStreamSynchronize(a.last_stream);   // reuse is ordered by the stream
    cudaGetLastError();
    a.off = 0;
    a.voff = 0;
    if (a.want > a.cap) arena_grow(a, a.want);
    a.last_stream = s;
    return CSB200_OK;
}

void arena_hint(size_t bytes)
{
    Arena &a = arena();
    if (bytes > a.want) a.want = bytes;
    if (a.depth == 1 && a.off == 0 && a.want > a.cap) {
        arena_grow(a, a.want);
        a.last_stream = stream();
    }
}

int ensure_device()
{
    static std::mutex mu;
    static bool tuned[64] = {false};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_error(CSB200_ERR_CUDA, "no usable CUDA device: %s (libcsparse_b200 has no CPU fallback)",
                         cudaGetErrorString(e));
    }
    if (dev >= 0 && dev < 64 && !tuned[dev]) {
        std::lock_guard<std::mutex> lk(mu);
        if (!tuned[dev]) {
            cudaMemPool_t pool;
            e = cudaDeviceGetDefaultMemPool(&pool, dev);
            if (e != cudaSuccess) {
                cudaGetLastError();
                return set_error(CSB200_ERR_CUDA, "no usable CUDA device: %s (libcsparse_b200 has no CPU fallback)",
                                 cudaGetErrorString(e));
            }
            unsigned long long keep = ~0ull;   // keep freed blocks cached in the pool
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            tuned[dev] = true;
        }
    }
    return CSB200_OK;
}

// implemented in transpose.cu / spmv.cu / spgemm.cu
int transpose_impl(const csb200_mat *A, bool values, csb200_mat **out);
int transpose_last_path();
int spmv_run(csb200_mat *AT, const double *d_x, double *d_y);
int spmv_build_plan(csb200_mat *AT);
void spmv_plan_free(SpmvPlan *pl);
int multiply_impl(csb200_mat *A, csb200_mat *B, csb200_mat **out, bool ordered);
int spmv_plan_kind(const SpmvPlan *pl);
int spmv_rows_align(csb200_mat *AT, int *align);
int spmv_run_rows(csb200_mat *AT, const double *d_x, double *d_y, int ra, int rb, cudaStream_t s);
int spmv_chunk_maxcol(csb200_mat *AT, int rows, int count, const int **out);
int halo_pull_launch(csb200_halo *h, cudaStream_t s);
int halo_acks_launch(csb200_halo *h, cudaStream_t s);
double *halo_window_ptr(csb200_halo *h);
long long halo_window_count(csb200_halo *h);

// copy streams and events of the chunked host pipeline of csb200_gaxpy, one set per thread
struct HostPipe {
    static constexpr int MAX_CHUNKS = 8;
    cudaStream_t h2d = nullptr, d2h = nullptr;
    cudaEvent_t start = nullptr, xdone = nullptr, end = nullptr, up[MAX_CHUNKS] = {}, done[MAX_CHUNKS] = {};
    int device = -1;
    int init()
    {
        int dev = 0;
        CSB_CUDA(cudaGetDevice(&dev));
        if (device == dev) return CSB200_OK;
        CSB_CUDA(cudaStreamCreateWithFlags(&h2d, cudaStreamNonBlocking));
        CSB_CUDA(cudaStreamCreateWithFlags(&d2h, cudaStreamNonBlocking));
        cudaEvent_t *all[] = {&start, &xdone, &end};
        for (cudaEvent_t *e : all) CSB_CUDA(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (int c = 0; c < MAX_CHUNKS; c++) {
            CSB_CUDA(cudaEventCreateWithFlags(&up[c], cudaEventDisableTiming));
            CSB_CUDA(cudaEventCreateWithFlags(&done[c], cudaEventDisableTiming));
        }
        device = dev;
        return CSB200_OK;
    }
};
static HostPipe &host_pipe()
{
    static thread_local HostPipe hp;
    return hp;
}

// p[0] == 0, p non-decreasing, 0 <= i < m
__global__ void k_validate(int m, int n, const csi *__restrict__ p, const csi *__restrict__ i, long long nnz, int *bad)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long j = t0; j < n; j += stride)
        if (p[j] > p[j + 1] || p[j] < 0) *bad = 1;
    if (t0 == 0 && p[0] != 0) *bad = 1;
    for (long long q = t0; q < nnz; q += stride) {
        const int r = i[q];
        if (r < 0 || r >= m) *bad = 2;
    }
}

__global__ void k_rebase(const csi *__restrict__ src, int count, int base, csi *__restrict__ dst)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < count) dst[k] = src[k] - base;
}

int mat_alloc(csi m, csi n, long long nnz, bool has_x, csb200_mat **out)
{
    csb200_mat *A = new (std::nothrow) csb200_mat();
    if (!A) return set_error(CSB200_ERR_NOMEM, "out of host memory");
    A->m = m; A->n = n; A->nnz = nnz;
    cudaGetDevice(&A->device);
    const size_t cap = (size_t)(nnz > 0 ? nnz : 1) + MAT_PAD;
    int st = dev_alloc(&A->p, (size_t)n + 1 + MAT_PAD);
    if (st == CSB200_OK) st = dev_alloc(&A->i, cap);
    if (st == CSB200_OK && has_x) st = dev_alloc(&A->x, cap);
    if (st != CSB200_OK) { csb200_mat_free(A); return st; }
    *out = A;
    return CSB200_OK;
}

static int ensure_csr(csb200_mat *A)
{
    if (A->csr) return CSB200_OK;
    if (!A->x) return set_error(CSB200_ERR_ARG, "cs_gaxpy: matrix has no values");
    CSB_TRY(transpose_impl(A, true, &A->csr));
    A->csr->forced_plan = A->forced_plan;
    return CSB200_OK;
}

}  // namespace csb

using namespace csb;

extern "C" {

int csb200_version(void) { return 100; }

const char *csb200_last_error(void) { return tls().err; }

int csb200_device_count(int *count)
{
    if (!count) return set_error(CSB200_ERR_ARG, "null argument");
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return set_error(CSB200_ERR_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    return CSB200_OK;
}

int csb200_set_device(int device)
{
    CSB_CUDA(cudaSetDevice(device));
    return CSB200_OK;
}

int csb200_set_stream(void *cuda_stream)
{
    tls().stream = (cudaStream_t)cuda_stream;
    return CSB200_OK;
}

int csb200_synchronize(void)
{
    CSB_CUDA(cudaStreamSynchronize(stream()));
    return CSB200_OK;
}

int csb200_sm_count(int *count)
{
    if (!count) return set_error(CSB200_ERR_ARG, "null argument");
    int dev = 0;
    CSB_CUDA(cudaGetDevice(&dev));
    CSB_CUDA(cudaDeviceGetAttribute(count, cudaDevAttrMultiProcessorCount, dev));
    return CSB200_OK;
}

int64_t csb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- cs_cumsum ---------------------------------------------------------------------
int csb200_cumsum_dev(csi *d_p, csi *d_c, csi n, int64_t *total)
{
    ArenaScope arena_scope;
    if (!d_p || !d_c || n < 0) return set_error(CSB200_ERR_ARG, "cs_cumsum: null array or n < 0");
    // the total is written by the kernel straight into pinned host memory (unified addressing): the
    // call is one memset, one launch and one synchronize -- a pageable 8-byte copy back cost ~15 us of
    // the 0.11 ms a scan of 2^24 counts took
    ThreadState &ts = tls();
    if (!ts.host_scalar) CSB_CUDA(cudaHostAlloc((void **)&ts.host_scalar, 64, cudaHostAllocPortable | cudaHostAllocMapped));
    CSB_TRY(launch_excl_scan(d_p, d_c, n, ts.host_scalar, nullptr));
    CSB_CUDA(cudaStreamSynchronize(stream()));
    const long long h = *reinterpret_cast<volatile long long *>(ts.host_scalar);
    if (total) *total = h;
    if (h > 0x7fffffffLL || h < -0x80000000LL)
        return set_error(CSB200_ERR_OVERFLOW, "cs_cumsum: total %lld does not fit int32", h);
    return CSB200_OK;
}

int csb200_cumsum(csi *p, csi *c, csi n, int64_t *total)
{
    ArenaScope arena_scope;
    if (!p || !c || n < 0) return set_error(CSB200_ERR_ARG, "cs_cumsum: null array or n < 0");
    DevBuf<csi> d_p, d_c;
    CSB_TRY(d_p.alloc((size_t)n + 1));
    CSB_TRY(d_c.alloc((size_t)n + 1));
    if (n > 0) CSB_CUDA(cudaMemcpyAsync(d_c.ptr, c, (size_t)n * sizeof(csi), cudaMemcpyHostToDevice, stream()));
    int st = csb200_cumsum_dev(d_p.ptr, d_c.ptr, n, total);
    if (st != CSB200_OK && st != CSB200_ERR_OVERFLOW) return st;
    CSB_CUDA(cudaMemcpyAsync(p, d_p.ptr, ((size_t)n + 1) * sizeof(csi), cudaMemcpyDeviceToHost, stream()));
    if (n > 0) CSB_CUDA(cudaMemcpyAsync(c, d_c.ptr, (size_t)n * sizeof(csi), cudaMemcpyDeviceToHost, stream()));
    CSB_CUDA(cudaStreamSynchronize(stream()));
    return st;
}

// ---- handles ---------------------------------------------------------------------------
int csb200_mat_upload(csi m, csi n, const csi *p, const csi *i, const double *x, int validate,
                      csb200_mat **out)
{
    ArenaScope arena_scope;
    if (!out) return set_error(CSB200_ERR_ARG, "null out");
    *out = nullptr;
    if (m < 0 || n < 0 || !p || (!i && p[n] > 0)) return set_error(CSB200_ERR_ARG, "mat_upload: bad arguments");
    const long long nnz = p[n];
    if (nnz < 0) return set_error(CSB200_ERR_INDEX, "mat_upload: p[n] < 0");
    csb200_mat *A = nullptr;
    CSB_TRY(mat_alloc(m, n, nnz, x != nullptr, &A));
    cudaStream_t s = stream();
    cudaError_t e = cudaMemcpyAsync(A->p, p, ((size_t)n + 1) * sizeof(csi), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(A->i, i, (size_t)nnz * sizeof(csi), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && nnz > 0 && x) e = cudaMemcpyAsync(A->x, x, (size_t)nnz * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && nnz == 0) {
        e = cudaMemsetAsync(A->i, 0, sizeof(csi), s);
        if (e == cudaSuccess && x) e = cudaMemsetAsync(A->x, 0, sizeof(double), s);
    }
    if (e != cudaSuccess) {
        csb200_mat_free(A);
        return set_error(CSB200_ERR_CUDA, "mat_upload copy: %s", cudaGetErrorString(e));
    }
    if (validate) {
        DevBuf<int> bad;
        int st = bad.alloc(1);
        if (st != CSB200_OK) { csb200_mat_free(A); return st; }
        cudaMemsetAsync(bad.ptr, 0, sizeof(int), s);
        const long long work = nnz > n ? nnz : n;
        const long long cap_blocks = (long long)sm_count() * 32;
        const int blocks = (int)(work / 256 + 1 < cap_blocks ? work / 256 + 1 : cap_blocks);
        k_validate<<<blocks, 256, 0, s>>>(m, n, A->p, A->i, nnz, bad.ptr);
        g_launches.fetch_add(1, std::memory_order_relaxed);
        int h = 0;
        e = cudaMemcpyAsync(&h, bad.ptr, sizeof(int), cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            csb200_mat_free(A);
            return set_error(CSB200_ERR_CUDA, "mat_upload validate: %s", cudaGetErrorString(e));
        }
        if (h) {
            csb200_mat_free(A);
            return set_error(CSB200_ERR_INDEX, h == 1 ? "column pointers are not a monotone sequence from 0"
                                                      : "row index outside [0, m)");
        }
    } else {
        // the host buffers may be reused by the caller as soon as we return
        e = cudaStreamSynchronize(s);
        if (e != cudaSuccess) {
            csb200_mat_free(A);
            return set_error(CSB200_ERR_CUDA, "mat_upload: %s", cudaGetErrorString(e));
        }
    }
    *out = A;
    return CSB200_OK;
}

int csb200_mat_from_dev(csi m, csi n, const csi *d_p, const csi *d_i, const double *d_x, csb200_mat **out)
{
    if (!out) return set_error(CSB200_ERR_ARG, "null out");
    *out = nullptr;
    if (m < 0 || n < 0 || !d_p) return set_error(CSB200_ERR_ARG, "mat_from_dev: bad arguments");
    csi nnz32 = 0;
    CSB_CUDA(cudaMemcpyAsync(&nnz32, d_p + n, sizeof(csi), cudaMemcpyDeviceToHost, stream()));
    CSB_CUDA(cudaStreamSynchronize(stream()));
    const long long nnz = nnz32;
    if (nnz < 0 || (nnz > 0 && !d_i)) return set_error(CSB200_ERR_ARG, "mat_from_dev: bad nnz / null i");
    csb200_mat *A = nullptr;
    CSB_TRY(mat_alloc(m, n, nnz, d_x != nullptr, &A));
    cudaStream_t s = stream();
    cudaError_t e = cudaMemcpyAsync(A->p, d_p, ((size_t)n + 1) * sizeof(csi), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(A->i, d_i, (size_t)nnz * sizeof(csi), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && nnz > 0 && d_x) e = cudaMemcpyAsync(A->x, d_x, (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) {
        csb200_mat_free(A);
        return set_error(CSB200_ERR_CUDA, "mat_from_dev copy: %s", cudaGetErrorString(e));
    }
    *out = A;
    return CSB200_OK;
}

int csb200_mat_dims(const csb200_mat *A, csi *m, csi *n, int64_t *nnz, int *has_values)
{
    if (!A) return set_error(CSB200_ERR_ARG, "null matrix");
    if (m) *m = A->m;
    if (n) *n = A->n;
    if (nnz) *nnz = A->nnz;
    if (has_values) *has_values = A->x != nullptr;
    return CSB200_OK;
}

int csb200_mat_download(const csb200_mat *A, csi *p, csi *i, double *x)
{
    if (!A) return set_error(CSB200_ERR_ARG, "null matrix");
    cudaStream_t s = stream();
    if (p) CSB_CUDA(cudaMemcpyAsync(p, A->p, ((size_t)A->n + 1) * sizeof(csi), cudaMemcpyDeviceToHost, s));
    if (i && A->nnz > 0) CSB_CUDA(cudaMemcpyAsync(i, A->i, (size_t)A->nnz * sizeof(csi), cudaMemcpyDeviceToHost, s));
    if (x && A->x && A->nnz > 0) CSB_CUDA(cudaMemcpyAsync(x, A->x, (size_t)A->nnz * sizeof(double), cudaMemcpyDeviceToHost, s));
    CSB_CUDA(cudaStreamSynchronize(s));
    return CSB200_OK;
}

int csb200_mat_dev_ptrs(const csb200_mat *A, csi **d_p, csi **d_i, double **d_x)
{
    if (!A) return set_error(CSB200_ERR_ARG, "null matrix");
    if (d_p) *d_p = A->p;
    if (d_i) *d_i = A->i;
    if (d_x) *d_x = A->x;
    return CSB200_OK;
}

int csb200_mat_col_slice(const csb200_mat *A, csi j0, csi j1, csb200_mat **out)
{
    if (!A || !out || j0 < 0 || j1 < j0 || j1 > A->n) return set_error(CSB200_ERR_ARG, "mat_col_slice: bad range");
    *out = nullptr;
    csi h[2] = {0, 0};
    CSB_CUDA(cudaMemcpyAsync(&h[0], A->p + j0, sizeof(csi), cudaMemcpyDeviceToHost, stream()));
    CSB_CUDA(cudaMemcpyAsync(&h[1], A->p + j1, sizeof(csi), cudaMemcpyDeviceToHost, stream()));
    CSB_CUDA(cudaStreamSynchronize(stream()));
    const long long nnz = (long long)h[1] - h[0];
    csb200_mat *S = nullptr;
    CSB_TRY(mat_alloc(A->m, j1 - j0, nnz, A->x != nullptr, &S));
    cudaStream_t s = stream();
    const int count = j1 - j0 + 1;
    k_rebase<<<ceil_div(count, 256), 256, 0, s>>>(A->p + j0, count, h[0], S->p);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && nnz > 0) e = cudaMemcpyAsync(S->i, A->i + h[0], (size_t)nnz * sizeof(csi), cudaMemcpyDeviceToDevice, s);
    if (e == cudaSuccess && nnz > 0 && A->x) e = cudaMemcpyAsync(S->x, A->x + h[0], (size_t)nnz * sizeof(double), cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) {
        csb200_mat_free(S);
        return set_error(CSB200_ERR_CUDA, "mat_col_slice: %s", cudaGetErrorString(e));
    }
    S->canon = A->canon;
    *out = S;
    return CSB200_OK;
}

int csb200_mat_free(csb200_mat *A)
{
    if (!A) return CSB200_OK;
    if (A->csr) csb200_mat_free(A->csr);
    if (A->plan) spmv_plan_free(A->plan);
    dev_free(A->p);
    dev_free(A->i);
    dev_free(A->x);
    dev_free(A->c32_blk);
    dev_free(A->c32_mask);
    dev_free(A->c32_len);
    dev_free(A->cls);
    dev_free(A->soa_x);
    delete A;
    return CSB200_OK;
}

// ---- cs_transpose ---------------------------------------------------------------------
int csb200_transpose(const csb200_mat *A, int values, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!A || !C) return set_error(CSB200_ERR_ARG, "cs_transpose: null argument");
    *C = nullptr;
    return transpose_impl(A, values != 0, C);
}

int csb200_transpose_force_path(int path)
{
    if (path < 0 || path > 4) return set_error(CSB200_ERR_ARG, "bad transpose path");
    tls().force_transpose = path;
    return CSB200_OK;
}

int csb200_transpose_last_path(void) { return transpose_last_path(); }

int csb200_transpose_host(csi m, csi n, const csi *Ap, const csi *Ai, const double *Ax,
                          csi *Cp, csi *Ci, double *Cx)
{
    ArenaScope arena_scope;
    if (!Cp || !Ci) return set_error(CSB200_ERR_ARG, "cs_transpose: null output");
    csb200_mat *A = nullptr, *C = nullptr;
    CSB_TRY(csb200_mat_upload(m, n, Ap, Ai, (Ax && Cx) ? Ax : nullptr, 1, &A));
    int st = transpose_impl(A, Cx != nullptr, &C);
    if (st == CSB200_OK) st = csb200_mat_download(C, Cp, Ci, Cx);
    csb200_mat_free(A);
    csb200_mat_free(C);
    return st;
}

// ---- cs_gaxpy -------------------------------------------------------------------------
int csb200_gaxpy_prepare(csb200_mat *A)
{
    ArenaScope arena_scope;
    if (!A) return set_error(CSB200_ERR_ARG, "cs_gaxpy: null matrix");
    CSB_TRY(ensure_csr(A));
    return spmv_build_plan(A->csr);
}

int csb200_gaxpy_plan(csb200_mat *A, int *kind)
{
    ArenaScope arena_scope;
    if (!A || !kind) return set_error(CSB200_ERR_ARG, "null argument");
    CSB_TRY(csb200_gaxpy_prepare(A));
    *kind = spmv_plan_kind(A->csr->plan);
    return CSB200_OK;
}

int csb200_gaxpy_force_plan(csb200_mat *A, int kind)
{
    if (!A || kind < 0 || kind > 4) return set_error(CSB200_ERR_ARG, "bad plan kind");
    A->forced_plan = kind;
    if (A->csr) A->csr->forced_plan = kind;
    return CSB200_OK;
}

int csb200_gaxpy_dev(csb200_mat *A, const double *d_x, double *d_y)
{
    ArenaScope arena_scope;
    if (!A || !d_x || !d_y) return set_error(CSB200_ERR_ARG, "cs_gaxpy: null argument");
    CSB_TRY(ensure_csr(A));
    return spmv_run(A->csr, d_x, d_y);
}

int csb200_gaxpy_t_dev(csb200_mat *AT, const double *d_x, double *d_y)
{
    ArenaScope arena_scope;
    if (!AT || !d_x || !d_y) return set_error(CSB200_ERR_ARG, "cs_gaxpy: null argument");
    return spmv_run(AT, d_x, d_y);
}

int csb200_gaxpy(csb200_mat *A, const double *x, double *y)
{
    ArenaScope arena_scope;
    if (!A || !x || !y) return set_error(CSB200_ERR_ARG, "cs_gaxpy: null argument");
    if (!A->x) return set_error(CSB200_ERR_ARG, "cs_gaxpy: matrix has no values");
    DevBuf<double> d_x, d_y;
    CSB_TRY(d_x.alloc((size_t)A->n));
    CSB_TRY(d_y.alloc((size_t)A->m));
    cudaStream_t s = stream();
    // Large row-stream matrices: y travels in row chunks on two copy streams, so the D2H of a
    // finished chunk overlaps the H2D of the next ones (PCIe is full duplex) and the SpMV of a
    // chunk starts as soon as its slice of y and the part of x it reads have landed.
    int align = 0;
    if (A->m >= (1 << 20)) {
        CSB_TRY(ensure_csr(A));
        CSB_TRY(spmv_rows_align(A->csr, &align));
    }
    if (align > 0) {
        HostPipe &hp = host_pipe();
        CSB_TRY(hp.init());
        const int m = A->m;
        int rows = (m + HostPipe::MAX_CHUNKS - 1) / HostPipe::MAX_CHUNKS;
        rows = ((rows + align - 1) / align) * align;
        const int nchunks = (m + rows - 1) / rows;
        const int *maxcol = nullptr;
        CSB_TRY(spmv_chunk_maxcol(A->csr, rows, nchunks, &maxcol));
        CSB_CUDA(cudaEventRecord(hp.start, s));                      // d_x / d_y exist, earlier work is done
        CSB_CUDA(cudaStreamWaitEvent(hp.h2d, hp.start, 0));
        // x goes up in pieces as well, each just before the first row chunk that reads it: a banded
        // matrix starts computing after 2/8 of x; an unstructured one needs all of x first, as before
        const long long nx = A->n;
        const long long xstep = ((nx + nchunks - 1) / nchunks + 1) & ~1LL;
        long long x_sent = 0;
        int c = 0;
        for (int ra = 0; ra < m; ra += rows, c++) {
            const int rb = ra + rows < m ? ra + rows : m;
            const size_t bytes = (size_t)(rb - ra) * sizeof(double);
            long long need = (long long)maxcol[c] + 1;               // x[0 .. need) must have landed
            for (int u = 0; u < c; u++) need = need > maxcol[u] + 1 ? need : maxcol[u] + 1;
            long long upto = ((need + xstep - 1) / xstep) * xstep;
            if (upto > nx) upto = nx;
            if (upto > x_sent) {
                CSB_CUDA(cudaMemcpyAsync(d_x.ptr + x_sent, x + x_sent, (size_t)(upto - x_sent) * sizeof(double),
                                         cudaMemcpyHostToDevice, hp.h2d));
                x_sent = upto;
            }
            CSB_CUDA(cudaMemcpyAsync(d_y.ptr + ra, y + ra, bytes, cudaMemcpyHostToDevice, hp.h2d));
            CSB_CUDA(cudaEventRecord(hp.up[c], hp.h2d));
            CSB_CUDA(cudaStreamWaitEvent(s, hp.up[c], 0));
            CSB_TRY(spmv_run_rows(A->csr, d_x.ptr, d_y.ptr, ra, rb, s));
            CSB_CUDA(cudaEventRecord(hp.done[c], s));
            CSB_CUDA(cudaStreamWaitEvent(hp.d2h, hp.done[c], 0));
            CSB_CUDA(cudaMemcpyAsync(y + ra, d_y.ptr + ra, bytes, cudaMemcpyDeviceToHost, hp.d2h));
        }
        CSB_CUDA(cudaEventRecord(hp.xdone, hp.h2d));
        CSB_CUDA(cudaEventRecord(hp.end, hp.d2h));
        CSB_CUDA(cudaStreamWaitEvent(s, hp.xdone, 0));               // the buffers are freed on s after the last copies
        CSB_CUDA(cudaStreamWaitEvent(s, hp.end, 0));
        CSB_CUDA(cudaStreamSynchronize(hp.d2h));
        return CSB200_OK;
    }
    if (A->n > 0) CSB_CUDA(cudaMemcpyAsync(d_x.ptr, x, (size_t)A->n * sizeof(double), cudaMemcpyHostToDevice, s));
    if (A->m > 0) CSB_CUDA(cudaMemcpyAsync(d_y.ptr, y, (size_t)A->m * sizeof(double), cudaMemcpyHostToDevice, s));
    CSB_TRY(csb200_gaxpy_dev(A, d_x.ptr, d_y.ptr));
    if (A->m > 0) CSB_CUDA(cudaMemcpyAsync(y, d_y.ptr, (size_t)A->m * sizeof(double), cudaMemcpyDeviceToHost, s));
    CSB_CUDA(cudaStreamSynchronize(s));
    return CSB200_OK;
}

// ---- cs_gaxpy on a row block of a sharded matrix, HOST x and y -------------------------------------
// The sharded step as the reference-facing call would see it: this rank's slice of x and of y live
// in (pinned) host memory.  The two ends of x travel first, so that the neighbours can pull their halo
// lines (k_halo_pull) while the rest of x and y are still on the bus; y then moves in row chunks on
// two copy streams exactly as in csb200_gaxpy -- the D2H of a finished chunk overlaps the H2D of the
// next ones, the SpMV of a chunk starts when the part of x it reads has landed.
int csb200_gaxpy_halo(csb200_mat *AT, csb200_halo *h, const double *x_own, int64_t own_off, int64_t own_len,
                      int64_t edge_lo, int64_t edge_hi, double *y, double *d_y_resident)
{
    ArenaScope arena_scope;
    if (!AT || !h || !x_own || (!y && !d_y_resident) || own_off < 0 || own_len < 0 || edge_lo < 0 || edge_hi < 0)
        return set_error(CSB200_ERR_ARG, "cs_gaxpy: bad arguments");
    if (!AT->x) return set_error(CSB200_ERR_ARG, "cs_gaxpy: matrix has no values");
    if (own_off + own_len > halo_window_count(h) || AT->m > halo_window_count(h))
        return set_error(CSB200_ERR_ARG, "cs_gaxpy: the x window is shorter than the block needs");
    if (edge_lo > own_len) edge_lo = own_len;
    if (edge_hi > own_len) edge_hi = own_len;
    const int m = AT->n;
    double *win = halo_window_ptr(h);
    double *d_own = win + own_off;
    // y: a host vector that travels up and down (d_y_resident == NULL), or a vector that stays in HBM
    // and accumulates there, of which the host gets a copy (y != NULL) after every step
    DevBuf<double> d_y_tmp;
    const bool y_up = d_y_resident == nullptr;
    if (y_up) CSB_TRY(d_y_tmp.alloc((size_t)(m > 0 ? m : 1)));
    struct { double *ptr; } d_y = {y_up ? d_y_tmp.ptr : d_y_resident};
    cudaStream_t s = stream();
    HostPipe &hp = host_pipe();
    CSB_TRY(hp.init());
    int align = 0;
    if (m >= (1 << 20)) CSB_TRY(spmv_rows_align(AT, &align));
    CSB_CUDA(cudaEventRecord(hp.start, s));                          // earlier work on s is done, d_y exists
    CSB_CUDA(cudaStreamWaitEvent(hp.h2d, hp.start, 0));
    // 1. the ends of my slice, for the neighbours
    if (edge_lo > 0) CSB_CUDA(cudaMemcpyAsync(d_own, x_own, (size_t)edge_lo * sizeof(double), cudaMemcpyHostToDevice, hp.h2d));
    if (edge_hi > 0)
        CSB_CUDA(cudaMemcpyAsync(d_own + own_len - edge_hi, x_own + own_len - edge_hi, (size_t)edge_hi * sizeof(double),
                                 cudaMemcpyHostToDevice, hp.h2d));
    CSB_CUDA(cudaEventRecord(hp.xdone, hp.h2d));
    CSB_CUDA(cudaStreamWaitEvent(s, hp.xdone, 0));
    CSB_TRY(halo_pull_launch(h, s));
    if (align > 0) {
        int rows = (m + HostPipe::MAX_CHUNKS - 1) / HostPipe::MAX_CHUNKS;
        rows = ((rows + align - 1) / align) * align;
        const int nchunks = (m + rows - 1) / rows;
        const int *maxcol = nullptr;
        CSB_TRY(spmv_chunk_maxcol(AT, rows, nchunks, &maxcol));
        const long long xstep = ((own_len + nchunks - 1) / nchunks + 1) & ~1LL;
        long long x_sent = 0;                                        // own coordinates
        int c = 0;
        for (int ra = 0; ra < m; ra += rows, c++) {
            const int rb = ra + rows < m ? ra + rows : m;
            const size_t bytes = (size_t)(rb - ra) * sizeof(double);
            long long need = 0;                                      // window coordinates: x[0 .. need) must be there
            for (int u = 0; u <= c; u++) need = need > (long long)maxcol[u] + 1 ? need : (long long)maxcol[u] + 1;
            long long upto = need - own_off;                         // the halo below own_off comes from the pull
            upto = ((upto + xstep - 1) / xstep) * xstep;
            if (upto > own_len) upto = own_len;
            if (upto > x_sent) {
                CSB_CUDA(cudaMemcpyAsync(d_own + x_sent, x_own + x_sent, (size_t)(upto - x_sent) * sizeof(double),
                                         cudaMemcpyHostToDevice, hp.h2d));
                x_sent = upto;
            }
            if (y_up) CSB_CUDA(cudaMemcpyAsync(d_y.ptr + ra, y + ra, bytes, cudaMemcpyHostToDevice, hp.h2d));
            CSB_CUDA(cudaEventRecord(hp.up[c], hp.h2d));
            CSB_CUDA(cudaStreamWaitEvent(s, hp.up[c], 0));
            CSB_TRY(spmv_run_rows(AT, win, d_y.ptr, ra, rb, s));
            CSB_CUDA(cudaEventRecord(hp.done[c], s));
            CSB_CUDA(cudaStreamWaitEvent(hp.d2h, hp.done[c], 0));
            if (y) CSB_CUDA(cudaMemcpyAsync(y + ra, d_y.ptr + ra, bytes, cudaMemcpyDeviceToHost, hp.d2h));
        }
    } else {
        if (own_len > 0) CSB_CUDA(cudaMemcpyAsync(d_own, x_own, (size_t)own_len * sizeof(double), cudaMemcpyHostToDevice, hp.h2d));
        if (m > 0 && y_up) CSB_CUDA(cudaMemcpyAsync(d_y.ptr, y, (size_t)m * sizeof(double), cudaMemcpyHostToDevice, hp.h2d));
        CSB_CUDA(cudaEventRecord(hp.up[0], hp.h2d));
        CSB_CUDA(cudaStreamWaitEvent(s, hp.up[0], 0));
        CSB_TRY(spmv_run(AT, win, d_y.ptr));
        CSB_CUDA(cudaEventRecord(hp.done[0], s));
        CSB_CUDA(cudaStreamWaitEvent(hp.d2h, hp.done[0], 0));
        if (m > 0 && y) CSB_CUDA(cudaMemcpyAsync(y, d_y.ptr, (size_t)m * sizeof(double), cudaMemcpyDeviceToHost, hp.d2h));
    }
    // the neighbours have pulled my lines: the next call may overwrite x
    CSB_TRY(halo_acks_launch(h, s));
    CSB_CUDA(cudaEventRecord(hp.end, hp.d2h));
    CSB_CUDA(cudaStreamWaitEvent(s, hp.end, 0));
    CSB_CUDA(cudaStreamSynchronize(s));
    return CSB200_OK;
}

int csb200_gaxpy_host(csi m, csi n, const csi *Ap, const csi *Ai, const double *Ax,
                      const double *x, double *y)
{
    ArenaScope arena_scope;
    if (!Ax) return set_error(CSB200_ERR_ARG, "cs_gaxpy: matrix has no values");
    if (!x || !y) return set_error(CSB200_ERR_ARG, "cs_gaxpy: null vector");
    csb200_mat *A = nullptr;
    CSB_TRY(csb200_mat_upload(m, n, Ap, Ai, Ax, 1, &A));
    int st = csb200_gaxpy(A, x, y);
    csb200_mat_free(A);
    return st;
}

// ---- cs_multiply ------------------------------------------------------------------------
int csb200_multiply(const csb200_mat *A, const csb200_mat *B, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!A || !B || !C) return set_error(CSB200_ERR_ARG, "cs_multiply: null argument");
    *C = nullptr;
    if (A->n != B->m) return set_error(CSB200_ERR_ARG, "cs_multiply: A.n != B.m");   // csparse.py:1618-1619
    return multiply_impl(const_cast<csb200_mat *>(A), const_cast<csb200_mat *>(B), C, false);
}

int csb200_multiply_ordered(const csb200_mat *A, const csb200_mat *B, csb200_mat **C)
{
    ArenaScope arena_scope;
    if (!A || !B || !C) return set_error(CSB200_ERR_ARG, "cs_multiply: null argument");
    *C = nullptr;
    if (A->n != B->m) return set_error(CSB200_ERR_ARG, "cs_multiply: A.n != B.m");
    return multiply_impl(const_cast<csb200_mat *>(A), const_cast<csb200_mat *>(B), C, true);
}

int csb200_multiply_force_path(int path)
{
    if (path < 0 || path > 6) return set_error(CSB200_ERR_ARG, "bad cs_multiply path");
    tls().multiply_ordered = path == 1;
    tls().multiply_blocked_version = (path == 2 || path == 3) ? path : 0;
    tls().multiply_templates = path == 4 ? 1 : path == 5 ? 2 : path == 6 ? 3 : 0;
    return CSB200_OK;
}

int64_t csb200_multiply_last_flops(void) { return tls().last_flops; }
int64_t csb200_multiply_last_templated(void) { return tls().last_templated; }

}  // extern "C"
