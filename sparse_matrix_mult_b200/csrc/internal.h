// internal.h -- host-side declarations shared by the .cu files of libspgemm_b200 (not part of the ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace sb {

// ---- bins ---------------------------------------------------------------------------------------
// symbolic bins, by the row's upper bound u = min(P_i, columns left)
enum { SYM_W64 = 0, SYM_W256, SYM_W1K, SYM_BITMAP, SYM_BINS };
// numeric bins, by the row's exact nnz
enum { NUM_W64 = 0, NUM_W256, NUM_W1K, NUM_RANK, NUM_BINS };

constexpr int kWarpCap64 = 48, kWarpCap256 = 192, kWarpCap1K = 768;   // <= 75 % load of the warp tables

struct LaunchCtx {
    cudaStream_t stream;
    int sm_count;
    int* launches;      // incremented once per kernel launch
    // side streams for kernels of one phase that are independent of each other (the cost-bin kernels): on small
    // inputs each of them is latency-bound with a handful of blocks, so they are run concurrently
    cudaStream_t aux[3];
    cudaEvent_t fork_ev;
    cudaEvent_t join_ev[3];
};

// ---- analysis.cu --------------------------------------------------------------------------------
// Per-row product counts of rows [row_begin, row_begin+nrows) of A against B, symbolic binning.
//   d_prod   : int64[nrows] or null
//   d_nnz    : int32[nrows]; rows with no product get 0 here (they enter no bin)
//   d_lists  : int32[SYM_BINS * nrows], bin b's rows (local ids) at d_lists + b*nrows
//   d_cursor : int32[SYM_BINS], zeroed by the caller; ends as the bin sizes
//   d_total  : unsigned long long, zeroed by the caller; ends as sum of products
cudaError_t launch_row_products(const LaunchCtx& lc, const Csr& A, const Csr& B, int row_begin, int nrows,
                                bool upper_only, int64_t* d_prod, int32_t* d_nnz, int32_t* d_lists,
                                int32_t* d_cursor, unsigned long long* d_total);

// d_flags (device int32[8]): [0] = 1 when every row of X has non-decreasing column indices, [3] = number of
// invalid entries (column out of range, indptr not monotone / not ending at nnz); [4] == [5] when every row is one
// run of consecutive columns; [1], [2] scratch.
cudaError_t launch_check_csr(const LaunchCtx& lc, const Csr& X, int64_t nnz, int32_t* d_flags);

// out[0] = 0, out[i+1] = out[i] + in[i]  (in int32[n], out OutT[n+1]); d_tmp holds >= 1025 int64.
cudaError_t launch_scan_i64(const LaunchCtx& lc, const int32_t* in, int64_t* out, int n, int64_t* d_tmp);
cudaError_t launch_scan_i32(const LaunchCtx& lc, const int32_t* in, int32_t* out, int n, int64_t* d_tmp);

// numeric binning by exact nnz
cudaError_t launch_bin_by_nnz(const LaunchCtx& lc, const int32_t* d_nnz, int nrows, int32_t* d_lists,
                              int32_t* d_cursor);

// CSR transpose: counts -> scan -> fill.  Rows of the transpose come out in arbitrary order.
cudaError_t launch_transpose_count(const LaunchCtx& lc, const Csr& X, int64_t nnz, int32_t* d_counts);
cudaError_t launch_transpose_fill(const LaunchCtx& lc, const Csr& X, int64_t nnz, const int32_t* t_ptr,
                                  int32_t* d_cursor, int32_t* t_idx, double* t_val);

// paneled transpose of rows [row_begin, row_end) of X (see analysis.cu): counts / t_ptr / cursor have np * X.cols (+1)
// entries, np = ceil((row_end - row_begin) / panel_w)
cudaError_t launch_transpose_count_panels(const LaunchCtx& lc, const Csr& X, int64_t nnz, int row_begin, int row_end,
                                          int panel_w, int32_t* d_counts);
cudaError_t launch_transpose_fill_panels(const LaunchCtx& lc, const Csr& X, int64_t nnz, int row_begin, int row_end,
                                         int panel_w, const int32_t* t_ptr, int32_t* d_cursor, uint32_t* t_pk,
                                         double* t_val);

// rows sorted by column (ascending or descending), in place, any row length; d_long_list: int32[rows + 1] scratch
cudaError_t launch_sort_rows(const LaunchCtx& lc, int rows, const int32_t* ptr, int32_t* idx, double* val,
                             int32_t* d_long_list, bool descending);
cudaError_t launch_narrow_indptr(const LaunchCtx& lc, const int64_t* in, int32_t* out, int n);

cudaError_t launch_add_const(const LaunchCtx& lc, int64_t* d_costs, int n, long long c);
// cost model of the triple product rows (see spgemm_b200_row_costs)
cudaError_t launch_triple_costs(const LaunchCtx& lc, const Csr& H, const Csr& Q, const Csr& Ht, bool upper_only,
                                bool q_runs, int np, int panel_w, int64_t* d_costs);

// ---- spgemm_sparse.cu ---------------------------------------------------------------------------
struct SparseJob {
    Csr A, B;
    int row_begin, nrows;
    bool upper_only;
    const int32_t* d_b_sorted;   // device flag
};
// Column bitmaps of heavy rows handed from the symbolic to the numeric phase (so the numeric rank kernel does not
// rebuild them with one more pass over the row's products).  bits == nullptr: nothing is kept.
struct SavedBitmaps {
    unsigned* bits;            // slots * words
    int32_t* slot_of_row;      // [nrows], -1 = not kept (preset by the caller)
    int32_t* counter;          // next free slot (zeroed by the caller)
    int slots, words;
};
int saved_bitmap_words(int cols);
cudaError_t launch_symbolic(const LaunchCtx& lc, const SparseJob& job, const int32_t* d_lists,
                            const int32_t* h_counts /* host, SYM_BINS */, int32_t* d_nnz, int32_t* d_work_counter,
                            const SavedBitmaps& saved);
cudaError_t launch_numeric(const LaunchCtx& lc, const SparseJob& job, const int32_t* d_lists,
                           const int32_t* h_counts /* host, NUM_BINS */, const int64_t* c_ptr, int32_t* c_idx,
                           double* c_val, int32_t* d_work_counter, const SavedBitmaps& saved);
cudaError_t sparse_kernels_configure();

// ---- spgemm_dense.cu ----------------------------------------------------------------------------
// mode: 0 = choose by products per output element, 1 = shared-memory tiles, 2 = block per row with L2 reductions
cudaError_t launch_dense(const LaunchCtx& lc, const Csr& A, const Csr& B, const int32_t* d_b_sorted,
                         bool upper_only, int row_begin, int nrows, double* d_c, int mode, double products_per_out);
cudaError_t launch_mirror(const LaunchCtx& lc, double* d_c, int n);
cudaError_t launch_symmetrize(const LaunchCtx& lc, double* d_c, int n);
cudaError_t dense_kernels_configure();

// ---- triple.cu ----------------------------------------------------------------------------------
// Column panels of C = H Q H^T (see k_triple_panels): np panels of panel_w columns starting at column k0.
struct TriplePlan {
    int k0, np, panel_w;
};
// An entry (r, c) of the paneled transpose is one 32-bit word: (r - first row of its panel) << 7 | (c & 127)
constexpr int kPanelColBits = 7;
constexpr uint32_t kPanelColMask = (1u << kPanelColBits) - 1u;
constexpr int kPanelMaxWidth = 1 << (32 - kPanelColBits);        // rows of H per panel the packing can address
TriplePlan triple_plan(int n, int row_begin, bool upper_only, int64_t h_nnz, int h_cols);
// t_ptr / t_pk / t_val: paneled transpose of rows [plan.k0, n) of H (launch_transpose_*_panels); q_runs: every row
// of Q is one run of consecutive, ascending columns (banded Q), which selects the lean kernel
cudaError_t launch_triple_panels(const LaunchCtx& lc, const Csr& H, const Csr& Q, bool q_runs, const int32_t* t_ptr,
                                 const uint32_t* t_pk, const double* t_val, const TriplePlan& plan, bool upper_only,
                                 int row_begin, int nrows, double* d_c,
                                 unsigned long long* d_counters /* [4], zeroed: P1, P2, ticket, spare */,
                                 int4* d_entry_meta /* nnz(H) entries of workspace for the banded-Q kernel, or null */);
cudaError_t triple_kernels_configure();

}  // namespace sb
