/*
 * oracle/spgemm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded restatement of the reference's CSR sparse x sparse
 * routines.  It exists only so that tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py can check (or time beside)
 * the CUDA path.  Nothing under sparse_matrix_mult_b200/ may link, load or
 * call it; the product path has no CPU fallback.
 *
 * Parity pinning: this file is checked (tests/test_oracle.py) against
 *   - the committed golden vectors in tests/golden/ that were produced by the
 *     unmodified reference Python wrapper + its shipped libsparse_x86_64.so
 *     (tests/golden/make_golden.py), and
 *   - when oracle/_ref/ is present, the reference binaries themselves
 *     (shipped serial .so for all five modes; from-source OpenMP build for
 *     dense_nosym / dense_sym / triple_product).
 *
 * The restated algorithm is that of the *shipped* revision (see SURVEY.md
 * section 0.3): Gustavson with a position-marker array initialised to -1 and
 * columns appended in first-touch order.  Today's src/ differs only by a
 * defective memory-pool rewrite; the loop structure cited below is the same.
 *
 * Interface: raw pointers and sizes only (no struct layouts), int32 CSR in,
 * int64 row pointers out so reduced R-MAT cases cannot overflow.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Result handle of the sparse-output routines. */
typedef struct {
    int64_t  nnz;
    int64_t  rows;
    int64_t *indptr;   /* rows + 1 */
    int32_t *indices;  /* nnz, first-touch order inside each row */
    double  *values;   /* nnz */
} oracle_csr;

void oracle_csr_free(oracle_csr *c)
{
    if (!c) return;
    free(c->indptr); free(c->indices); free(c->values); free(c);
}

/*
 * C = A*B (upper_only == 0) or the col >= row part of it (upper_only != 0).
 *
 * Follows  src/sparsework.cpp:56-129  (sparsework_nosym hot loop: marker test
 * :73-77, append :105-110, per-row count :116, marker reset :120-128) and
 * src/sparsework.cpp:201-280 (sparsework_sym; the only difference is the
 * `col >= row` filter at :217).  The per-thread partitioning of
 * src/sparse_sparse_sparse.cpp:228-249 and the stitch at :265-291 reduce, for
 * one thread, to running rows 0..m-1 in order and prefix-summing the per-row
 * counts -- which is what happens here.  Entries whose value cancels to 0.0
 * stay in the structure (the reference never prunes).
 */
oracle_csr *oracle_spgemm_csr(int m, int k, int n,
                              const int32_t *a_ptr, const int32_t *a_idx, const double *a_val,
                              const int32_t *b_ptr, const int32_t *b_idx, const double *b_val,
                              int upper_only)
{
    (void)k;
    oracle_csr *c = (oracle_csr *)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->rows = m;
    c->indptr = (int64_t *)calloc((size_t)m + 1, sizeof(int64_t));
    int64_t cap = 1024;
    c->indices = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
    c->values = (double *)malloc((size_t)cap * sizeof(double));
    int64_t *where = (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
    if (!c->indptr || !c->indices || !c->values || !where) { free(where); oracle_csr_free(c); return NULL; }
    for (int j = 0; j < n; ++j) where[j] = -1;

    int64_t fill = 0;
    for (int i = 0; i < m; ++i) {
        const int64_t row_begin = fill;
        for (int32_t p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
            const double av = a_val[p];
            const int32_t j = a_idx[p];
            for (int32_t q = b_ptr[j]; q < b_ptr[j + 1]; ++q) {
                const int32_t col = b_idx[q];
                if (upper_only && col < i) continue;          /* sparsework.cpp:217 */
                if (where[col] >= row_begin) {                 /* seen in this row  */
                    c->values[where[col]] += av * b_val[q];
                } else {                                       /* first touch       */
                    if (fill == cap) {
                        cap *= 2;
                        int32_t *ni = (int32_t *)realloc(c->indices, (size_t)cap * sizeof(int32_t));
                        double *nv = (double *)realloc(c->values, (size_t)cap * sizeof(double));
                        if (!ni || !nv) { free(where); if (ni) c->indices = ni; if (nv) c->values = nv; oracle_csr_free(c); return NULL; }
                        c->indices = ni; c->values = nv;
                    }
                    c->indices[fill] = col;
                    c->values[fill] = av * b_val[q];
                    where[col] = fill;
                    ++fill;
                }
            }
        }
        c->indptr[i + 1] = fill;
        /* where[] entries < row_begin of the next row are stale by construction,
           so no reset pass is needed (the reference resets to -1, same effect). */
    }
    free(where);
    c->nnz = fill;
    return c;
}

/*
 * Multi-threaded variant of oracle_spgemm_csr for the CPU baseline of bench.py only ("port", many cores): the
 * reference's design -- contiguous row blocks per OpenMP thread, private result buffers, serial stitch
 * (src/sparse_sparse_sparse.cpp:188-197, 228-249, 265-291) -- restated so that it works (today's src/ does not,
 * SURVEY.md 0.3).  Rows are computed exactly as in the serial routine, so the result is bit-identical to it
 * (tests/test_oracle.py).  Blocks are many and handed out dynamically: power-law inputs unbalance an even split.
 * Compiled without OpenMP this is the serial routine over one block.
 */
oracle_csr *oracle_spgemm_csr_omp(int m, int k, int n,
                                  const int32_t *a_ptr, const int32_t *a_idx, const double *a_val,
                                  const int32_t *b_ptr, const int32_t *b_idx, const double *b_val,
                                  int upper_only, int blocks)
{
    (void)k;
    if (blocks < 1) blocks = 1;
    if (blocks > m) blocks = m > 0 ? m : 1;
    oracle_csr *c = (oracle_csr *)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->rows = m;
    c->indptr = (int64_t *)calloc((size_t)m + 1, sizeof(int64_t));
    int32_t **part_idx = (int32_t **)calloc((size_t)blocks, sizeof *part_idx);
    double **part_val = (double **)calloc((size_t)blocks, sizeof *part_val);
    int64_t *part_nnz = (int64_t *)calloc((size_t)blocks, sizeof *part_nnz);
    int failed = !c->indptr || !part_idx || !part_val || !part_nnz;
#ifdef _OPENMP
#pragma omp parallel if (!failed)
#endif
    {
        int64_t *where = failed ? NULL : (int64_t *)malloc(((size_t)n + 1) * sizeof(int64_t));
        if (where) for (int j = 0; j < n; ++j) where[j] = -1;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
        for (int blk = 0; blk < blocks; ++blk) {
            if (!where) { failed = 1; continue; }
            const int r0 = (int)((int64_t)m * blk / blocks), r1 = (int)((int64_t)m * (blk + 1) / blocks);
            int64_t cap = 1024, fill = 0;
            int32_t *ci = (int32_t *)malloc((size_t)cap * sizeof(int32_t));
            double *cv = (double *)malloc((size_t)cap * sizeof(double));
            /* positions are block-local: offset them so that stale marks of earlier blocks never look fresh */
            for (int i = r0; i < r1 && ci && cv; ++i) {
                const int64_t row_begin = fill;
                for (int32_t p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
                    const double av = a_val[p];
                    const int32_t j = a_idx[p];
                    for (int32_t q = b_ptr[j]; q < b_ptr[j + 1]; ++q) {
                        const int32_t col = b_idx[q];
                        if (upper_only && col < i) continue;
                        if (where[col] >= row_begin && where[col] < fill && ci[where[col]] == col) {
                            cv[where[col]] += av * b_val[q];
                        } else {
                            if (fill == cap) {
                                cap *= 2;
                                int32_t *ni = (int32_t *)realloc(ci, (size_t)cap * sizeof(int32_t));
                                double *nv = (double *)realloc(cv, (size_t)cap * sizeof(double));
                                if (!ni || !nv) { if (ni) ci = ni; if (nv) cv = nv; failed = 1; break; }
                                ci = ni; cv = nv;
                            }
                            ci[fill] = col;
                            cv[fill] = av * b_val[q];
                            where[col] = fill;
                            ++fill;
                        }
                    }
                }
                c->indptr[i + 1] = fill - row_begin;          /* per-row count, prefix-summed below */
            }
            if (!ci || !cv) failed = 1;
            part_idx[blk] = ci; part_val[blk] = cv; part_nnz[blk] = fill;
        }
        free(where);
    }
    if (!failed) {
        for (int i = 0; i < m; ++i) c->indptr[i + 1] += c->indptr[i];
        c->nnz = c->indptr[m];
        c->indices = (int32_t *)malloc((size_t)(c->nnz ? c->nnz : 1) * sizeof(int32_t));
        c->values = (double *)malloc((size_t)(c->nnz ? c->nnz : 1) * sizeof(double));
        failed = !c->indices || !c->values;
    }
    if (!failed) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1)
#endif
        for (int blk = 0; blk < blocks; ++blk) {
            const int r0 = (int)((int64_t)m * blk / blocks);
            memcpy(c->indices + c->indptr[r0], part_idx[blk], (size_t)part_nnz[blk] * sizeof(int32_t));
            memcpy(c->values + c->indptr[r0], part_val[blk], (size_t)part_nnz[blk] * sizeof(double));
        }
    }
    if (part_idx) for (int b = 0; b < blocks; ++b) free(part_idx[b]);
    if (part_val) for (int b = 0; b < blocks; ++b) free(part_val[b]);
    free(part_idx); free(part_val); free(part_nnz);
    if (failed) { oracle_csr_free(c); return NULL; }
    return c;
}

/*
 * Dense C (m x n, row-major, caller-allocated, overwritten) = A*B.
 * Follows src/sparse_sparse_dense.cpp:108-130 (dense_nosym) and :40-73
 * (dense_sym, `i <= col` filter at :59; the lower triangle stays 0 because the
 * mirror code at :63-70 is commented out in the reference).
 */
void oracle_spgemm_dense(int m, int k, int n,
                         const int32_t *a_ptr, const int32_t *a_idx, const double *a_val,
                         const int32_t *b_ptr, const int32_t *b_idx, const double *b_val,
                         int upper_only, double *c)
{
    (void)k;
    memset(c, 0, (size_t)m * (size_t)n * sizeof(double));
    for (int i = 0; i < m; ++i) {
        double *row = c + (size_t)i * (size_t)n;
        for (int32_t p = a_ptr[i]; p < a_ptr[i + 1]; ++p) {
            const double av = a_val[p];
            const int32_t j = a_idx[p];
            for (int32_t q = b_ptr[j]; q < b_ptr[j + 1]; ++q) {
                const int32_t col = b_idx[q];
                if (upper_only && col < i) continue;
                row[col] += av * b_val[q];
            }
        }
    }
}

/*
 * Dense C (n x n, caller-allocated, overwritten) = H*Q*H^T, H is n x k, Q is k x k.
 * Follows src/sparse_sparse_dense.cpp:185-220: phase 1 (:187-198) expands
 * t = H[i,:]*Q into a dense length-k scratch row, phase 2 (:201-216) takes the
 * sparse dot of t with every row r >= i of H (r >= 0 when full != 0).
 * full != 0 reproduces the reference's behaviour literally, including its
 * double write at :213-215: every off-diagonal entry receives T[i,r] + T[r,i]
 * (SURVEY.md section 0.5), i.e. C = T + T^T - diag(T).
 */
void oracle_triple_product(int n, int k,
                           const int32_t *h_ptr, const int32_t *h_idx, const double *h_val,
                           const int32_t *q_ptr, const int32_t *q_idx, const double *q_val,
                           int full, double *c)
{
    memset(c, 0, (size_t)n * (size_t)n * sizeof(double));
    double *t = (double *)calloc((size_t)k > 0 ? (size_t)k : 1, sizeof(double));
    if (!t) return;
    for (int i = 0; i < n; ++i) {
        for (int32_t p = h_ptr[i]; p < h_ptr[i + 1]; ++p) {
            const int32_t j = h_idx[p];
            const double hv = h_val[p];
            for (int32_t q = q_ptr[j]; q < q_ptr[j + 1]; ++q)
                t[q_idx[q]] += hv * q_val[q];
        }
        for (int r = full ? 0 : i; r < n; ++r) {
            double dot = 0.0;
            for (int32_t p = h_ptr[r]; p < h_ptr[r + 1]; ++p)
                dot += t[h_idx[p]] * h_val[p];
            c[(size_t)i * n + r] += dot;
            if (full && r != i) c[(size_t)r * n + i] += dot;
        }
        memset(t, 0, (size_t)k * sizeof(double));
    }
    free(t);
}

/*
 * Row split of src/workdivision.cpp:16-89 (`limits`): parts = min(parts, rows)
 * (:26-29); the first rows % parts partitions get one extra row (:46-52,71-77);
 * out[p] = first row, out[p + parts] = last row (inclusive) (:63-65).
 * Returns the number of partitions actually produced, 0 on bad input (the
 * reference calls exit(0) there, :19-23).
 */
int oracle_limits(int rows, int parts, int32_t *out)
{
    if (parts <= 0 || rows <= 0) return 0;
    if (parts > rows) parts = rows;
    const int base = rows / parts, extra = rows % parts;
    int next = 0;
    for (int p = 0; p < parts; ++p) {
        const int len = base + (p < extra ? 1 : 0);
        out[p] = next;
        out[p + parts] = next + len - 1;
        next += len;
    }
    return parts;
}

/* Number of intermediate products P = sum over nnz (i,j) of A of nnz(B[j,:]);
   SURVEY.md section 8 notation.  Used by bench.py for flops = 2*P. */
int64_t oracle_count_products(int m, const int32_t *a_ptr, const int32_t *a_idx, const int32_t *b_ptr)
{
    int64_t total = 0;
    for (int i = 0; i < m; ++i)
        for (int32_t p = a_ptr[i]; p < a_ptr[i + 1]; ++p)
            total += b_ptr[a_idx[p] + 1] - b_ptr[a_idx[p]];
    return total;
}

#ifdef __cplusplus
}
#endif
