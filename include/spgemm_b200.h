/*
 * spgemm_b200.h -- C ABI of libspgemm_b200.so, the B200 (sm_100a) CUDA replacement for the native
 * library behind sparse_matrix_mult.sparse_matrix_multiply.
 *
 * The reference crosses Python -> C through ctypes with struct pointers
 *   (/root/reference/include/functions.h:43-84, called from
 *    /root/reference/sparse_matrix_mult/matrix_ops.py:195,333,346,348,351,360,362,365)
 * and the struct layouts on the two sides disagree (matrix_def.h:17-31 `size_t` vs
 * matrix_ops.py:26-33,44-48 `c_int`).  This ABI therefore passes plain pointers and sizes only:
 * int32 CSR index arrays, float64 values, caller-owned output buffers, `int` status returns
 * (0 = ok) with a thread-local message behind spgemm_b200_last_error().
 *
 * Which reference interface each entry point replaces is stated on the entry point.
 * There is no CPU fallback: every compute call fails with SPGEMM_B200_ERR_CUDA when no sm_100 device
 * is usable.
 */
#ifndef SPGEMM_B200_H
#define SPGEMM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SPGEMM_B200_API __attribute__((visibility("default")))
#else
#define SPGEMM_B200_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SPGEMM_B200_OK            0
#define SPGEMM_B200_ERR_ARG       1   /* bad argument (null pointer, negative size, dimension mismatch) */
#define SPGEMM_B200_ERR_CUDA      2   /* CUDA runtime error; see spgemm_b200_last_error() */
#define SPGEMM_B200_ERR_OVERFLOW  3   /* result does not fit the requested index width */
#define SPGEMM_B200_ERR_STATE     4   /* call order (e.g. copy of a freed result) */

/* triple-product output modes (argument `mode` below) */
#define SPGEMM_B200_TRIPLE_UPPER      0  /* triu(H Q H^T), zeros below: reference compute_full_matrix=0 */
#define SPGEMM_B200_TRIPLE_REF_FULL   1  /* T + T^T - diag(T): what reference compute_full_matrix=1 returns
                                            (src/sparse_sparse_dense.cpp:201-215 writes both halves of a full loop) */
#define SPGEMM_B200_TRIPLE_MIRROR     2  /* upper triangle computed once, mirrored below: true symmetric result */

/* Opaque handles. */
typedef struct spgemm_b200_mat spgemm_b200_mat;        /* a CSR matrix resident in HBM            */
typedef struct spgemm_b200_result spgemm_b200_result;  /* a sparse product C resident in HBM      */
typedef struct spgemm_b200_multi_result spgemm_b200_multi_result;  /* a sparse product in row blocks on N GPUs */

/* Thread safety.  The library keeps per-call state (stats, staging buffers) in one context per CUDA device.
   Every entry point locks the context it works on for its whole duration and makes that context's device
   current for the call (the caller's current device is restored on return), so:
     - concurrent calls from several threads are safe; calls on the same device are serialised;
     - handles may be used from any thread; a handle remembers its device;
     - spgemm_b200_get_stats / _last_error describe the last call OF THE CALLING THREAD only if no other thread
       called in between (stats are per device, the error string is per thread);
     - spgemm_b200_shutdown must not run concurrently with other calls.
   (The reference C library is stateless: every call allocates and frees its own scratch,
   src/sparse_sparse_sparse.cpp:199-291.) */

/* Per-call timing/size record of the LAST compute call on the default device's context.
   All times are CUDA-event milliseconds on the library stream. */
typedef struct {
    double  ms_h2d;        /* host -> device copies of the operands                                   */
    double  ms_analysis;   /* row-product count + binning (+ sortedness check, transpose for triple)   */
    double  ms_symbolic;   /* sparse output: symbolic phase + scan                                    */
    double  ms_numeric;    /* sparse output: numeric phase;  dense/triple: the one compute kernel      */
    double  ms_post;       /* mirror / symmetrise kernels                                              */
    double  ms_d2h;        /* device -> host copy of the result                                        */
    double  ms_total;      /* first event to last event                                                */
    int64_t products;      /* P = number of intermediate products (flops = 2P); triple: P1 + P2        */
    int64_t nnz_c;         /* sparse output: nnz(C); dense/triple: rows*cols written                   */
    int64_t bytes_min;     /* algorithmic (compulsory) bytes of the compute phase, SURVEY.md 8(d)      */
    int32_t launches;      /* kernels launched by this library during the call                         */
    int32_t device;        /* CUDA device ordinal used                                                 */
    int64_t bytes_h2d;     /* bytes copied host -> device by the call                                  */
    int64_t bytes_d2h;     /* bytes copied device -> host by the call (upper trapezoids only for the
                              symmetric dense modes: the zeros below the diagonal are written by the host) */
} spgemm_b200_stats;

/* ---- library / device ------------------------------------------------------------------------- */

/* Library version string, e.g. "spgemm_b200 0.1 (sm_100a)". */
SPGEMM_B200_API const char *spgemm_b200_version(void);

/* Message of the last failing call on this thread ("" if none). */
SPGEMM_B200_API const char *spgemm_b200_last_error(void);

/* Number of visible CUDA devices (0 when there is no driver/GPU; never fails). */
SPGEMM_B200_API int spgemm_b200_device_count(void);

/* Bind the calling process to `device` (default 0 when never called): creates the stream, the
   stream-ordered memory pool configuration and the timing events.  Replaces the reference's implicit
   OpenMP team setup (omp_get_max_threads(), src/sparse_sparse_sparse.cpp:188-197). */
SPGEMM_B200_API int spgemm_b200_init(int device);

/* Release every cached device/pinned buffer and the streams of every device context. */
SPGEMM_B200_API void spgemm_b200_shutdown(void);

/* Give cached memory back: the library allocates from a PRIVATE stream-ordered pool per device (the device's
   default pool, which other libraries share, is never reconfigured) that keeps up to SPGEMM_B200_POOL_KEEP_GB
   (default 32) of freed workspaces for reuse, and caches the pinned buffers of the most recent results
   (SPGEMM_B200_PINNED_CACHE_GB, default 4, or the last result alone when it is larger).  This call trims
   every pool to `keep_bytes` and drops the pinned cache.  Replaces the free() calls of destroy_sparsemat /
   destroy_darray (src/memfunctions.cpp:22-65) as the point where memory returns to the system. */
SPGEMM_B200_API int spgemm_b200_trim(size_t keep_bytes);

/* Timing/size record of the last compute call. */
SPGEMM_B200_API int spgemm_b200_get_stats(spgemm_b200_stats *out);

/* Pinned host memory from a size-bucketed cache (so NumPy results can be written by DMA at PCIe speed).
   Replaces create_darray / destroy_darray (functions.h:54,45) as the owner of dense result storage. */
SPGEMM_B200_API void *spgemm_b200_host_alloc(size_t bytes);
SPGEMM_B200_API void  spgemm_b200_host_free(void *p);

/* ---- host-buffer entry points: what matrix_ops.py binds ---------------------------------------- */

/* C = A*B as CSR kept on the device; `upper_only` keeps col >= row only.
   Replaces sparse_nosym / sparse_sym (functions.h:76,80; src/sparse_sparse_sparse.cpp:172-299,41-155),
   sparsework_nosym / sparsework_sym (functions.h:61,65; src/sparsework.cpp:12-149,156-300) and
   limits (functions.h:57; src/workdivision.cpp:16-89).  A is m x k, B is k x n.
   Columns inside each row of C come out sorted ascending; numerically cancelled entries stay. */
SPGEMM_B200_API int spgemm_b200_csr(int m, int k, int n,
                    const int32_t *a_indptr, const int32_t *a_indices, const double *a_values,
                    const int32_t *b_indptr, const int32_t *b_indices, const double *b_values,
                    int upper_only, spgemm_b200_result **out);

/* nnz / shape of a result. */
SPGEMM_B200_API int64_t spgemm_b200_result_nnz(const spgemm_b200_result *r);
SPGEMM_B200_API int     spgemm_b200_result_rows(const spgemm_b200_result *r);
SPGEMM_B200_API int     spgemm_b200_result_cols(const spgemm_b200_result *r);

/* Copy a result to caller-owned host arrays: indptr has rows+1 entries of int32 (index64 == 0; fails with
   ERR_OVERFLOW when nnz >= 2^31) or int64 (index64 != 0); indices int32[nnz]; values double[nnz].
   Replaces the array reads of sparsemat_to_csr (matrix_ops.py:205-228). */
SPGEMM_B200_API int spgemm_b200_result_copy(const spgemm_b200_result *r, void *indptr, int index64,
                            int32_t *indices, double *values);

/* Free a result.  Replaces destroy_sparsemat (functions.h:43; matrix_ops.py:351). */
SPGEMM_B200_API void spgemm_b200_result_free(spgemm_b200_result *r);

/* Dense C (m x n row-major float64, host, caller-owned, fully overwritten) = A*B.
   upper_only: keep col >= row, zeros below (reference dense_sym).  mirror (needs upper_only, m == n): copy
   the upper triangle below the diagonal on the device before the copy out.
   Replaces dense_nosym / dense_sym (functions.h:72,69; src/sparse_sparse_dense.cpp:79-131,13-74). */
SPGEMM_B200_API int spgemm_b200_dense(int m, int k, int n,
                      const int32_t *a_indptr, const int32_t *a_indices, const double *a_values,
                      const int32_t *b_indptr, const int32_t *b_indices, const double *b_values,
                      int upper_only, int mirror, double *c_host);

/* Dense C (n x n row-major float64, host, caller-owned, fully overwritten) = H*Q*H^T with H n x k and
   Q k x k, fused (H*Q is never materialised).  `mode` is one of SPGEMM_B200_TRIPLE_*.
   Replaces triple_product (functions.h:84; src/sparse_sparse_dense.cpp:141-249). */
SPGEMM_B200_API int spgemm_b200_triple(int n, int k,
                       const int32_t *h_indptr, const int32_t *h_indices, const double *h_values,
                       const int32_t *q_indptr, const int32_t *q_indices, const double *q_values,
                       int mode, double *c_host);

/* ---- device-resident entry points (operands already in HBM; used for handle reuse, the
        kernel-only leg of bench.py and for row-sharded multi-GPU runs) --------------------------- */

/* Upload a host CSR to the device (copies).  Replaces create_sparsemat + memmove
   (functions.h:51; matrix_ops.py:187-202). */
SPGEMM_B200_API int spgemm_b200_mat_upload(int rows, int cols, int64_t nnz,
                           const int32_t *indptr, const int32_t *indices, const double *values,
                           spgemm_b200_mat **out);

/* Wrap device arrays the caller owns (e.g. torch tensors); nothing is copied or freed. */
SPGEMM_B200_API int spgemm_b200_mat_wrap(int rows, int cols, int64_t nnz,
                         const int32_t *d_indptr, const int32_t *d_indices, const double *d_values,
                         spgemm_b200_mat **out);

/* Build the CSR of X^T on the device; every row of the result is sorted by DESCENDING column (what the
   triple product's upper-triangle contraction wants). */
SPGEMM_B200_API int spgemm_b200_mat_transpose(const spgemm_b200_mat *x, spgemm_b200_mat **out);

/* Device-side canonicalisation (the reference coerces on the host,
   sparse_matrix_mult/matrix_ops.py:307-310): sort every row of x by ascending column (duplicates stay; the
   accumulators sum them).  In place for matrices the library uploaded; borrowed (wrapped) arrays are left
   alone and a sorted shadow copy is cached on the handle.  The compute entry points do this automatically for
   an unsorted right operand.  spgemm_b200_mat_is_sorted: 1 / 0, or -1 on error (runs the validation pass). */
SPGEMM_B200_API int spgemm_b200_mat_sort(spgemm_b200_mat *x);
SPGEMM_B200_API int spgemm_b200_mat_is_sorted(const spgemm_b200_mat *x);

/* Keep (enable != 0) or drop the paneled transpose that spgemm_b200_triple_dev builds from this matrix when it is the
   H of a triple product: a caller that multiplies with the same resident H repeatedly -- the iterations of an
   inversion -- then pays for the transpose once.  The reference has no counterpart: it rebuilds
   nothing because it never transposes (it dots t with every row of H, src/sparse_sparse_dense.cpp:201-216). */
SPGEMM_B200_API int spgemm_b200_mat_cache_transpose(spgemm_b200_mat *x, int enable);

SPGEMM_B200_API void spgemm_b200_mat_free(spgemm_b200_mat *m);

/* Rows [row_begin, row_end) of C = A*B as a device-resident CSR result (row_end < 0 means all rows).
   The result has row_end-row_begin rows. */
SPGEMM_B200_API int spgemm_b200_csr_dev(const spgemm_b200_mat *a, const spgemm_b200_mat *b, int upper_only,
                        int row_begin, int row_end, spgemm_b200_result **out);

/* Device pointers of a result (valid until result_free): indptr is int64[rows+1]. */
SPGEMM_B200_API int spgemm_b200_result_device_ptrs(const spgemm_b200_result *r, const int64_t **d_indptr,
                                   const int32_t **d_indices, const double **d_values);

/* Rows [row_begin, row_end) of the dense product into d_c (device, (row_end-row_begin) x n, row-major).
   upper_only uses the GLOBAL row number for the col >= row test. */
SPGEMM_B200_API int spgemm_b200_dense_dev(const spgemm_b200_mat *a, const spgemm_b200_mat *b, int upper_only,
                          int row_begin, int row_end, double *d_c);

/* Rows [row_begin, row_end) of H*Q*H^T into d_c (device, (row_end-row_begin) x n).  ht may be NULL (built
   internally).  upper_only as above.  Mirroring/symmetrising needs the whole matrix and is done by
   spgemm_b200_mirror_dev / spgemm_b200_symmetrize_dev on the gathered result. */
SPGEMM_B200_API int spgemm_b200_triple_dev(const spgemm_b200_mat *h, const spgemm_b200_mat *q, const spgemm_b200_mat *ht,
                           int upper_only, int row_begin, int row_end, double *d_c);

/* In place on an n x n device matrix: C[j,i] = C[i,j] for j > i. */
SPGEMM_B200_API int spgemm_b200_mirror_dev(double *d_c, int n);
/* In place on an n x n device matrix: C = C + C^T - diag(C). */
SPGEMM_B200_API int spgemm_b200_symmetrize_dev(double *d_c, int n);

/* Per-row cost of rows of A against B (int64[rows(A)], device) and their total: the flop-counting pass that replaces
   limits() (src/workdivision.cpp:16-89).  Sparse / dense output: the intermediate-product count P_i of the row, plus
   0.3 * dense_cols when the output row is a dense row of dense_cols doubles written in full (pass 0 for sparse output).
   Triple product (q != NULL; a = H, q = Q, b = H^T as built by spgemm_b200_mat_transpose):
   cost_i = a P1_i panels(i) + b P2_i (n - i)/n + c n, see k_triple_costs.  Either output may be NULL. */
SPGEMM_B200_API int spgemm_b200_row_costs(const spgemm_b200_mat *a, const spgemm_b200_mat *b, const spgemm_b200_mat *q,
                          int upper_only, int dense_cols, int64_t *d_costs, int64_t *total_host);

/* Flop-balanced contiguous row partition: bounds[parts+1] (host) with bounds[0] = 0, bounds[parts] = rows,
   chosen so every part carries ~total/parts of d_costs.  Replaces limits() for multi-GPU sharding. */
SPGEMM_B200_API int spgemm_b200_partition(const int64_t *d_costs, int rows, int parts, int32_t *bounds_host);

/* The same for the triple product, where the block that starts at row r also pays for the paneled transpose of rows
   r.. of H: its cost is tail_coeff * (entries of H from row r on) + the sum of its rows' costs, and the partition
   minimises the largest block cost.  indptr_host: H's row pointers on the host; tail_coeff in the units of
   spgemm_b200_row_costs (SPGEMM_B200_TRIPLE_TAIL_COEFF below is the fitted value). */
#define SPGEMM_B200_TRIPLE_TAIL_COEFF 8.6
SPGEMM_B200_API int spgemm_b200_partition_tail(const int64_t *d_costs, const int32_t *indptr_host, double tail_coeff,
                               int rows, int parts, int32_t *bounds_host);

/* Raw device buffers from the library's stream-ordered pool (for callers without their own allocator),
   and plain copies on the library stream (synchronous on return). */
SPGEMM_B200_API void *spgemm_b200_device_alloc(size_t bytes);
SPGEMM_B200_API void  spgemm_b200_device_free(void *d_ptr);
SPGEMM_B200_API int   spgemm_b200_copy_to_host(void *host_dst, const void *d_src, size_t bytes);
SPGEMM_B200_API int   spgemm_b200_copy_to_device(void *d_dst, const void *host_src, size_t bytes);
/* n x n device matrix whose strictly lower triangle is zero -> host: only the upper trapezoids cross PCIe, the
   rest is zeroed by host threads (what spgemm_b200_dense / _triple do for their symmetric modes); synchronous */
SPGEMM_B200_API int   spgemm_b200_copy_upper_to_host(double *host_dst, const double *d_src, int n);
/* device -> device on the library stream, asynchronous */
SPGEMM_B200_API int   spgemm_b200_copy_on_device(void *d_dst, const void *d_src, size_t bytes);

/* ---- peer memory (one process per GPU on one NVLink/NVSwitch box) -------------------------------------
   A rank allocates its result buffer with spgemm_b200_shared_alloc, exports it, and the other ranks map it
   with spgemm_b200_ipc_open: the row-range entry points above then WRITE THEIR ROWS STRAIGHT INTO THAT BUFFER
   over NVLink (the kernels only see a pointer), which fuses the gather-to-rank-0 into the compute kernel. */
#define SPGEMM_B200_IPC_HANDLE_BYTES 64
SPGEMM_B200_API void *spgemm_b200_shared_alloc(size_t bytes);                 /* cudaMalloc: exportable   */
SPGEMM_B200_API void  spgemm_b200_shared_free(void *d_ptr);
SPGEMM_B200_API int   spgemm_b200_ipc_export(const void *d_ptr, unsigned char *handle /* 64 bytes */);
SPGEMM_B200_API int   spgemm_b200_ipc_open(const unsigned char *handle /* 64 bytes */, void **d_ptr);
SPGEMM_B200_API int   spgemm_b200_ipc_close(void *d_ptr);

/* ---- single-process multi-GPU (N GPUs behind one call) --------------------------------------------------
   The reference sizes its OpenMP team inside the call (src/sparse_sparse_sparse.cpp:188-197) and hands every
   thread a row range from limits() (src/workdivision.cpp:16-89).  Here: one host thread per GPU inside the call
   (devices 0..n_gpus-1), rows split by the flop-balanced partition, every GPU loads the operands over its own
   PCIe link and writes its row block straight into the caller's host result (symmetric dense modes: upper
   trapezoids only).  Same argument meaning as spgemm_b200_dense / _triple (upper mode) / _csr. */
SPGEMM_B200_API int spgemm_b200_multi_dense(int n_gpus, int m, int k, int n,
                            const int32_t *a_indptr, const int32_t *a_indices, const double *a_values,
                            const int32_t *b_indptr, const int32_t *b_indices, const double *b_values,
                            int upper_only, double *c_host);
SPGEMM_B200_API int spgemm_b200_multi_triple(int n_gpus, int n, int k,
                             const int32_t *h_indptr, const int32_t *h_indices, const double *h_values,
                             const int32_t *q_indptr, const int32_t *q_indices, const double *q_values,
                             double *c_host);
SPGEMM_B200_API int spgemm_b200_multi_csr(int n_gpus, int m, int k, int n,
                          const int32_t *a_indptr, const int32_t *a_indices, const double *a_values,
                          const int32_t *b_indptr, const int32_t *b_indices, const double *b_values,
                          int upper_only, spgemm_b200_multi_result **out);
SPGEMM_B200_API int64_t spgemm_b200_multi_result_nnz(const spgemm_b200_multi_result *r);
/* every GPU copies its block to its final offset of the caller's arrays (parallel stitch; replaces
   src/sparse_sparse_sparse.cpp:265-291) */
SPGEMM_B200_API int  spgemm_b200_multi_result_copy(const spgemm_b200_multi_result *r, void *indptr, int index64,
                                   int32_t *indices, double *values);
SPGEMM_B200_API void spgemm_b200_multi_result_free(spgemm_b200_multi_result *r);
/* diagnostics of the last multi-GPU call: the row bounds (returns their count, n_gpus + 1) and the per-GPU stats */
SPGEMM_B200_API int spgemm_b200_multi_last_bounds(int32_t *bounds, int capacity);
SPGEMM_B200_API int spgemm_b200_multi_last_stats(int part, spgemm_b200_stats *out);

/* Make the library launch on `stream` (a cudaStream_t; NULL restores the library's own stream; pass
   cudaStreamLegacy, i.e. (cudaStream_t)0x1, to name the legacy default stream). */
SPGEMM_B200_API int spgemm_b200_set_stream(void *stream);

/* CUDA-event stopwatch on the library stream: start records an event, stop records another, waits for it and
   returns the milliseconds in between (what bench.py brackets each step with). */
SPGEMM_B200_API int spgemm_b200_timer_start(void);
SPGEMM_B200_API int spgemm_b200_timer_stop(double *ms);

/* Overwrite a scratch buffer larger than the L2 cache (evicts operands between timed iterations). */
SPGEMM_B200_API int spgemm_b200_flush_l2(void);

/* Block until the library stream is idle. */
SPGEMM_B200_API int spgemm_b200_synchronize(void);

#ifdef __cplusplus
}
#endif

#endif /* SPGEMM_B200_H */
