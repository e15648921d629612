/*
 * csparse_b200.h -- C ABI of libcsparse_b200.so
 *
 * B200-native (sm_100a) replacement for the data-parallel hot path of
 * rwl/CSparse.py: cs_cumsum, cs_transpose, cs_gaxpy, cs_multiply.  The
 * reference is pure Python and has no FFI layer: its boundary is
 * the module namespace.  These entry points are what a ctypes binding of that
 * namespace binds; each cites the reference function it
 * replaces as csparse.py:line.
 *
 * Conventions
 *   - plain C: pointers + sizes, int32 indices (csi), IEEE binary64 values.
 *   - every function returns a csb200_status; CSB200_ERR_ARG corresponds to
 *     the reference's sentinel returns (None / False / -1).
 *   - "host" entry points take caller-owned host buffers and do the H2D/D2H
 *     copies themselves; "_dev" entry points take device pointers and are
 *     asynchronous on the current stream (csb200_set_stream).
 *   - csb200_mat is an opaque, immutable, device-resident CSC matrix
 *     (the reference's `cs` object with nz == -1, csparse.py:37-54).
 *   - no CPU fallback: without a CUDA device every compute call fails with
 *     CSB200_ERR_CUDA.
 *   - threads and streams: the library keeps its state per thread (stream, last error, path
 *     switches, a workspace block that call temporaries are carved from).  Different threads may
 *     work on different handles concurrently.  A handle -- with the caches it builds on first use
 *     (CSR view, gaxpy plan and its scratch, pattern classes, entry-major copy) -- is bound to ONE
 *     thread and ONE stream at a time: use it, and free it, from the thread / stream that last
 *     used it, or synchronise in between (csb200_synchronize).  Calls that take host buffers
 *     return after the stream has drained; "_dev" calls and calls on handles only are
 *     asynchronous.
 */
#ifndef CSPARSE_B200_H
#define CSPARSE_B200_H

#include <stdint.h>

#if defined(__GNUC__)
#define CSB200_API __attribute__((visibility("default")))
#else
#define CSB200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef int32_t csi;
typedef struct csb200_mat csb200_mat;
typedef struct csb200_halo csb200_halo;

typedef enum {
    CSB200_OK = 0,
    CSB200_ERR_ARG = 1,       /* bad argument: reference returns None / False / -1 */
    CSB200_ERR_CUDA = 2,      /* CUDA runtime / driver failure (no device, OOM, ...) */
    CSB200_ERR_OVERFLOW = 3,  /* a count does not fit int32 */
    CSB200_ERR_INDEX = 4,     /* row index outside [0, m) or p not monotone */
    CSB200_ERR_NOMEM = 5
} csb200_status;

/* ---- library state ------------------------------------------------------ */
CSB200_API int csb200_version(void);
CSB200_API const char *csb200_last_error(void);              /* thread-local message of the last failure */
CSB200_API int csb200_device_count(int *count);
CSB200_API int csb200_set_device(int device);                /* cudaSetDevice for the calling thread */
CSB200_API int csb200_set_stream(void *cuda_stream);         /* stream used by this thread's calls; NULL = legacy default */
CSB200_API int csb200_synchronize(void);                     /* cudaStreamSynchronize on that stream */
CSB200_API int csb200_sm_count(int *count);
/* number of kernels this library launched since load (all threads); bench.py's gpu_launches */
CSB200_API int64_t csb200_launch_count(void);

/* ---- cs_cumsum (csparse.py:767-784) ------------------------------------- */
/* p[0..n] = exclusive prefix sum of c[0..n-1]; c[0..n-1] <- p[0..n-1];
 * *total = sum(c) as int64.  Only p[0..n] and c[0..n-1] are touched.
 * Single-pass decoupled look-back scan. */
CSB200_API int csb200_cumsum(csi *p, csi *c, csi n, int64_t *total);          /* host buffers */
CSB200_API int csb200_cumsum_dev(csi *d_p, csi *d_c, csi n, int64_t *total);  /* device buffers; synchronises to return total */

/* ---- matrix handles (the cs object, csparse.py:37-54) -------------------- */
/* nnz is p[n]; i/x may be longer than nnz (tails ignored, as in the reference).
 * x == NULL => pattern-only matrix (A.x is None).  validate != 0 checks that p
 * is monotone from 0 and every row index is in [0, m). */
CSB200_API int csb200_mat_upload(csi m, csi n, const csi *p, const csi *i, const double *x,
                      int validate, csb200_mat **out);
/* copy device arrays (d_p has n+1 entries) into a new handle */
CSB200_API int csb200_mat_from_dev(csi m, csi n, const csi *d_p, const csi *d_i, const double *d_x,
                        csb200_mat **out);
CSB200_API int csb200_mat_dims(const csb200_mat *A, csi *m, csi *n, int64_t *nnz, int *has_values);
CSB200_API int csb200_mat_download(const csb200_mat *A, csi *p, csi *i, double *x);   /* p: n+1, i/x: nnz */
CSB200_API int csb200_mat_dev_ptrs(const csb200_mat *A, csi **d_p, csi **d_i, double **d_x);
/* columns [j0, j1) as a new matrix (used to shard B / the CSR view across GPUs) */
CSB200_API int csb200_mat_col_slice(const csb200_mat *A, csi j0, csi j1, csb200_mat **out);
CSB200_API int csb200_mat_free(csb200_mat *A);

/* ---- cs_transpose (csparse.py:2292-2315) --------------------------------- */
/* C = A' as CSC (= CSR view of A): stable counting sort by row index.  Result
 * p, i, x are bit-identical to the reference's.  C has values iff values != 0
 * and A has values. */
CSB200_API int csb200_transpose(const csb200_mat *A, int values, csb200_mat **C);
/* which algorithm later transposes use: 0 = automatic (one-pass mirror lookup for square matrices
 * with sorted duplicate-free columns and a symmetric pattern -- verified by the kernel itself --,
 * else the two-level bucket sort, or the stable radix sort when power-law rows overflow the
 * buckets), 1 = always the radix sort, 2 = automatic without the mirror path, 3 = like 2 with the
 * bucket sort's partition and sort phases interleaved in L2-sized slabs (measured slower than two
 * whole passes), 4 = like 2 with partition and sort in one persistent launch in which a bucket is
 * sorted by the CTA that completes it (short rows only; measured slower too); for tests and benchmarks */
CSB200_API int csb200_transpose_force_path(int path);
/* the path the calling thread's last transpose took: 1 mirror, 2 bucket sort, 3 radix sort,
 * 0 trivial (empty matrix) */
CSB200_API int csb200_transpose_last_path(void);
/* one-shot form on host buffers: Cp has m+1 slots, Ci/Cx nnz slots (Cx may be NULL) */
CSB200_API int csb200_transpose_host(csi m, csi n, const csi *Ap, const csi *Ai, const double *Ax,
                          csi *Cp, csi *Ci, double *Cx);

/* ---- cs_gaxpy (csparse.py:1199-1213) -------------------------------------- */
/* y[0..m) += A * x[0..n).  Row-parallel, atomic-free SpMV over the CSR view of
 * A, which is built once per handle by csb200_transpose and cached. */
CSB200_API int csb200_gaxpy(csb200_mat *A, const double *x, double *y);           /* host x, y */
CSB200_API int csb200_gaxpy_dev(csb200_mat *A, const double *d_x, double *d_y);   /* device x, y; async */
/* y[0..AT.n) += AT' * x[0..AT.m): the same kernel applied to an explicit CSR
 * view (AT is the CSC of A').  Used by the row-block sharded multi-GPU path. */
CSB200_API int csb200_gaxpy_t_dev(csb200_mat *AT, const double *d_x, double *d_y);
/* one-shot form, everything on the host (matrix upload + CSR build + SpMV) */
CSB200_API int csb200_gaxpy_host(csi m, csi n, const csi *Ap, const csi *Ai, const double *Ax,
                      const double *x, double *y);
/* build (and cache) the CSR view now instead of on first use */
CSB200_API int csb200_gaxpy_prepare(csb200_mat *A);
/* which kernel the cached plan uses: 1 = row-stream (TMA-staged), 2 = merge-path,
 * 3 = row-stream with plain loads (kept for A/B measurements) */
CSB200_API int csb200_gaxpy_plan(csb200_mat *A, int *kind);
/* force a plan (0 = automatic); for tests and benchmarks */
CSB200_API int csb200_gaxpy_force_plan(csb200_mat *A, int kind);

/* ---- cs_gaxpy across GPUs: row blocks of the CSR view, x halos over NVLink ----
 * The reference has no distributed code; these are the in-library form of its cs_gaxpy
 * (csparse.py:1199-1213) for a matrix whose rows are split across one process per GPU.  A halo
 * object owns the rank's x window [lo halo | own slice | hi halo] (device memory) and a small
 * flag block; its IPC handles are exchanged once (128 bytes per rank, by whatever transport the
 * host has -- torch.distributed in csparse_cuda/dist.py), after which every step is ONE kernel
 * launch per rank: the persistent SpMV pulls the halo lines from the neighbours' windows through
 * peer-mapped pointers, runs the interior row blocks meanwhile and the edge row blocks last. */
CSB200_API int csb200_halo_create(int64_t count, csb200_halo **out);       /* window of `count` doubles, zeroed */
CSB200_API int csb200_halo_window(csb200_halo *h, double **d_window);
CSB200_API int csb200_halo_export(csb200_halo *h, void *handles);          /* 128 bytes out */
/* side 0 = the rank below, 1 = the rank above: window[local_first .. +count) is pulled from the
 * peer's window[peer_first .. +count) every step */
CSB200_API int csb200_halo_connect(csb200_halo *h, int side, const void *peer_handles, int64_t peer_first,
                        int64_t count, int64_t local_first);
/* the same for a peer living in this process (tests; one process driving several GPUs) */
CSB200_API int csb200_halo_connect_local(csb200_halo *h, int side, csb200_halo *peer, int64_t peer_first,
                              int64_t count, int64_t local_first);
/* y[0..AT.n) += AT' * window; top_rows / bot_rows = rows at the two ends of the block that read
 * halo entries (0, 0 with no neighbours).  Every rank must call it the same number of times. */
CSB200_API int csb200_gaxpy_halo_dev(csb200_mat *AT, csb200_halo *h, double *d_y, csi top_rows, csi bot_rows);
/* the same step on HOST vectors: x_own (own_len doubles) is this rank's slice of x, stored at offset
 * own_off of the window; edge_lo / edge_hi = how many doubles at the two ends of the slice the
 * neighbours read (they travel first).  d_y_resident == NULL: y (AT.n host doubles) is read and
 * written.  d_y_resident != NULL: y accumulates in that device vector and the host vector y, if not
 * NULL, receives a copy of the result.  Chunked duplex copies as in csb200_gaxpy.  Returns when y is
 * back in host memory and the neighbours have pulled their lines. */
CSB200_API int csb200_gaxpy_halo(csb200_mat *AT, csb200_halo *h, const double *x_own, int64_t own_off, int64_t own_len,
                                 int64_t edge_lo, int64_t edge_hi, double *y, double *d_y_resident);
CSB200_API int csb200_halo_status(csb200_halo *h, int *timed_out);         /* 1: a neighbour never showed up (2 s) */
CSB200_API int csb200_halo_free(csb200_halo *h);

/* ---- cs_multiply (csparse.py:1608-1642, cs_scatter :1961-1989) ------------ */
/* C = A*B: symbolic per-column count (shared-memory hash sets, dense spill for
 * large columns), exclusive scan, numeric fill.  Structural zeros are kept.
 * Columns of C come out in the reference's discovery order.  C has values iff
 * both A and B have. */
CSB200_API int csb200_multiply(const csb200_mat *A, const csb200_mat *B, csb200_mat **C);
/* cs_multiply with the rows of every column of C in the reference's discovery order (what the
 * Python layer uses for host `cs` operands, so that p, i, x equal the reference's bit for bit) */
CSB200_API int csb200_multiply_ordered(const csb200_mat *A, const csb200_mat *B, csb200_mat **C);
/* Order of the rows inside C's columns for csb200_multiply: 0 = automatic -- columns whose pattern fits the blocked
 * numeric kernel come out block by block (32-row blocks in discovery order, ascending inside a
 * block), the rest in the reference's discovery order; 1 = always the reference's discovery order
 * (then p, i, x are bit-identical to cs_multiply on canonical inputs).  The set of rows and every
 * value are the same either way.  2 / 3 = automatic with the first / second version of the
 * blocked numeric kernel, for A/B measurements and tests.  4 = automatic without the pattern-class
 * templates, 5 = templates tried at any size (automatic: from 16384 columns), 6 = like 5 with one warp
 * per column only (without the kernel that runs 32 columns of a class in lock step).  Columns formed from a
 * class template (matrices of translation-invariant operators: one symbolic pass per class of
 * columns instead of one per column) are always in the reference's discovery order. */
CSB200_API int csb200_multiply_force_path(int path);
/* columns the last csb200_multiply on this thread formed from pattern-class templates */
CSB200_API int64_t csb200_multiply_last_templated(void);
/* number of multiply-adds of the last csb200_multiply on this thread */
CSB200_API int64_t csb200_multiply_last_flops(void);

/* ---- the callers and data formats either side of the hot path ---- */
/* cs_add (csparse.py:163-192): C = alpha*A + beta*B as [A B] * [alpha I; beta I] on the SpGEMM
 * kernels: column order and rounding are the reference's.  C has values iff A and B have.
 * The handle holds exactly nnz(C) entries (the reference over-allocates nnz(A)+nnz(B)). */
CSB200_API int csb200_add(csb200_mat *A, csb200_mat *B, double alpha, double beta, csb200_mat **C);
/* cs_add's algorithm: 0 = automatic (sorted duplicate-free operands: binary-search merge kernels,
 * otherwise the SpGEMM kernels), 1 = always the SpGEMM kernels; for tests */
CSB200_API int csb200_add_force_path(int path);
/* cs_norm (csparse.py:1647-1663): largest column sum of |x|, entries added in storage order */
CSB200_API int csb200_norm(const csb200_mat *A, double *norm);
/* cs_compress (csparse.py:647-672): triplets (Ti, Tj, Tx or NULL; nz of them, any order,
 * duplicates allowed) -> CSC by a stable radix sort on the column index; entries of a column
 * keep their input order.  Host and device-buffer forms. */
CSB200_API int csb200_compress(csi m, csi n, csi nz, const csi *Ti, const csi *Tj, const double *Tx, csb200_mat **C);
CSB200_API int csb200_compress_dev(csi m, csi n, csi nz, const csi *d_Ti, const csi *d_Tj, const double *d_Tx,
                        csb200_mat **C);
/* cs_dupl (csparse.py:1035-1063): duplicates summed into their first occurrence, as A * I on the
 * SpGEMM kernels; out of place (handles are immutable).  A must have values. */
CSB200_API int csb200_dupl(csb200_mat *A, csb200_mat **C);
/* cs_fkeep (csparse.py:1172-1196) with a fixed predicate: 0 keep aij != 0 (cs_dropzeros :1024),
 * 1 keep |aij| > tol (cs_droptol :1007), 2 keep i != j (csparse_test.py Dropdiag, cs_amd's _cs_diag
 * :207-211), 3 keep i <= j, 4 keep the entries of columns with at most tol entries (cs_amd's
 * dense-column drop, :236-249).
 * Order inside the columns is kept; out of place. */
CSB200_API int csb200_fkeep(const csb200_mat *A, int predicate, double tol, csb200_mat **C);
/* cs_permute (csparse.py:1666-1693): C = P A Q; pinv (m entries) / q (n entries) are host arrays
 * or NULL for the identity */
CSB200_API int csb200_permute(const csb200_mat *A, const csi *pinv, const csi *q, int values, csb200_mat **C);
/* cs_symperm (csparse.py:2220-2255): C = P A P' using the upper triangular part of a symmetric A;
 * stable radix sort on max(pinv[i], pinv[j]) */
CSB200_API int csb200_symperm(const csb200_mat *A, const csi *pinv, int values, csb200_mat **C);

#ifdef __cplusplus
}
#endif
#endif /* CSPARSE_B200_H */
