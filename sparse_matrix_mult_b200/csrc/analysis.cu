// analysis.cu -- the passes that replace the reference's row partitioner (src/workdivision.cpp:16-89):
// per-row intermediate-product counts, cost binning, scans, CSR transpose, sortedness check.
#include <cstdio>
#include <cstdlib>

#include "internal.h"

namespace sb {

#define SB_LAUNCH_CHECK(lc)            \
    do {                               \
        ++*(lc).launches;              \
        cudaError_t e_ = cudaGetLastError(); \
        if (e_ != cudaSuccess) return e_;    \
    } while (0)

// ---------------------------------------------------------------------------------------------------
// Row products + symbolic binning.  Eight lanes per row of A: the gathers A.idx[p] -> B.ptr[j], B.ptr[j+1]
// of a row are issued together; rows are short on average (10-200 entries in the BASELINE configs).
__global__ void __launch_bounds__(256)
k_row_products(Csr A, const int32_t* __restrict__ b_ptr, int b_cols, int row_begin, int nrows, int upper_only,
               int64_t* __restrict__ prod, int32_t* __restrict__ nnz, int32_t* __restrict__ lists,
               int32_t* __restrict__ cursor, unsigned long long* __restrict__ total) {
    __shared__ int s_cnt[SYM_BINS], s_base[SYM_BINS];
    __shared__ unsigned long long s_total;
    if (threadIdx.x < SYM_BINS) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0) s_total = 0;
    __syncthreads();

    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    const int gl = threadIdx.x & 7;
    long long sum = 0;
    int i = 0;
    if (r < nrows) {
        i = row_begin + r;
        const int s = __ldg(A.ptr + i), e = __ldg(A.ptr + i + 1);
#pragma unroll 4
        for (int p = s + gl; p < e; p += 8) {
            const int j = __ldg(A.idx + p);
            sum += __ldg(b_ptr + j + 1) - __ldg(b_ptr + j);
        }
    }
    sum += __shfl_xor_sync(FULL, sum, 1);
    sum += __shfl_xor_sync(FULL, sum, 2);
    sum += __shfl_xor_sync(FULL, sum, 4);

    int bin = -1;
    if (r < nrows && gl == 0) {
        if (prod) prod[r] = sum;
        long long cap = upper_only ? (long long)b_cols - i : (long long)b_cols;
        long long u = sum < cap ? sum : cap;
        if (u <= 0) {
            nnz[r] = 0;
        } else {
            bin = u <= kWarpCap64 ? SYM_W64 : u <= kWarpCap256 ? SYM_W256 : u <= kWarpCap1K ? SYM_W1K : SYM_BITMAP;
        }
        if (sum) atomicAdd(&s_total, (unsigned long long)sum);
    }
    int local = 0;
    if (bin >= 0) local = atomicAdd(&s_cnt[bin], 1);
    __syncthreads();
    if (threadIdx.x < SYM_BINS) {
        const int c = s_cnt[threadIdx.x];
        s_base[threadIdx.x] = c ? atomicAdd(cursor + threadIdx.x, c) : 0;
    }
    if (threadIdx.x == 0 && s_total) atomicAdd(total, s_total);
    __syncthreads();
    if (bin >= 0) lists[(size_t)bin * nrows + s_base[bin] + local] = r;
}

cudaError_t launch_row_products(const LaunchCtx& lc, const Csr& A, const Csr& B, int row_begin, int nrows,
                                bool upper_only, int64_t* d_prod, int32_t* d_nnz, int32_t* d_lists,
                                int32_t* d_cursor, unsigned long long* d_total) {
    if (nrows <= 0) return cudaSuccess;
    const int threads = 256, rows_per_block = threads / 8;
    const int blocks = (nrows + rows_per_block - 1) / rows_per_block;
    k_row_products<<<blocks, threads, 0, lc.stream>>>(A, B.ptr, B.cols, row_begin, nrows, upper_only ? 1 : 0, d_prod,
                                                      d_nnz, d_lists, d_cursor, d_total);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------
// Validation + sortedness of a CSR in one pass over idx and ptr (an out-of-range index would become an
// out-of-bounds device atomic).  A CSR has sorted rows iff every descent idx[q] > idx[q+1] sits on a row boundary.
// flags[1] = descents anywhere, flags[2] = descents on row boundaries, flags[3] = invalid entries (column index
// outside [0, cols), indptr not monotone / outside [0, nnz], indptr[0] != 0, indptr[rows] != nnz);
// flags[4] = pairs with idx[q+1] != idx[q] + 1 anywhere, flags[5] = such pairs on row boundaries: every row is ONE
// RUN of consecutive columns (a banded matrix) iff the two are equal;
// flags[0] = rows sorted (set by k_sorted_flag).
__global__ void __launch_bounds__(256)
k_check_csr(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, int rows, int cols, int64_t nnz,
            int64_t work, int32_t* __restrict__ flags) {
    // grid-stride over `work` items; item t covers entries [4t, 4t + 4) (one 128-bit load when aligned, plus the first
    // entry of the next quad) and row t.  Counts are summed per thread, then per block, and reach the five global
    // counters with at most one atomic per block each (a banded matrix has a "gap" on every row boundary: one atomic
    // per warp to the same address took 0.7 ms for the 65 M entries of cfg 5's Q).
    int any = 0, edge = 0, bad = 0, gap = 0, gap_edge = 0;
    const bool vec = (reinterpret_cast<uintptr_t>(idx) & 15) == 0;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < work; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q0 = t * 4;
        if (q0 < nnz) {
            int c[5];
            if (q0 + 4 <= nnz && vec) {
                const int4 v = __ldg(reinterpret_cast<const int4*>(idx) + t);
                c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) c[u] = q0 + u < nnz ? __ldg(idx + q0 + u) : 0x7fffffff;
            }
            c[4] = q0 + 4 < nnz ? __ldg(idx + q0 + 4) : 0x7fffffff;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (q0 + u < nnz) {
                    bad += c[u] < 0 || c[u] >= cols;
                    if (q0 + u + 1 < nnz) {
                        any += c[u] > c[u + 1];
                        gap += c[u + 1] != c[u] + 1;
                    }
                }
            }
        }
        if (t < rows) {
            const int s = __ldg(ptr + t), e = __ldg(ptr + t + 1);
            if (s < 0 || e < s || (int64_t)e > nnz) ++bad;
            else if (e > s && e < nnz) {
                const int last = __ldg(idx + e - 1), next = __ldg(idx + e);
                edge += last > next;
                gap_edge += next != last + 1;
            }
            if (t == 0 && s != 0) ++bad;
            if (t == rows - 1 && (int64_t)e != nnz) ++bad;
        }
    }
    __shared__ int s_sum[5];
    if (threadIdx.x < 5) s_sum[threadIdx.x] = 0;
    __syncthreads();
    any = warp_sum(any); edge = warp_sum(edge); bad = warp_sum(bad); gap = warp_sum(gap); gap_edge = warp_sum(gap_edge);
    if (lane_id() == 0) {
        if (any) atomicAdd(&s_sum[0], any);
        if (edge) atomicAdd(&s_sum[1], edge);
        if (bad) atomicAdd(&s_sum[2], bad);
        if (gap) atomicAdd(&s_sum[3], gap);
        if (gap_edge) atomicAdd(&s_sum[4], gap_edge);
    }
    __syncthreads();
    if (threadIdx.x < 5 && s_sum[threadIdx.x]) atomicAdd(flags + 1 + threadIdx.x, s_sum[threadIdx.x]);
}
__global__ void k_sorted_flag(int32_t* __restrict__ flags) {
    flags[0] = flags[1] == flags[2] ? 1 : 0;
}

cudaError_t launch_check_csr(const LaunchCtx& lc, const Csr& X, int64_t nnz, int32_t* d_flags) {
    cudaError_t e = cudaMemsetAsync(d_flags, 0, 8 * sizeof(int32_t), lc.stream);
    if (e != cudaSuccess) return e;
    const int64_t quads = (nnz + 3) / 4;
    const int64_t n = quads > X.rows ? quads : X.rows;
    if (n > 0) {
        const int threads = 256;
        int64_t blocks = (n + threads - 1) / threads;
        if (blocks > (int64_t)lc.sm_count * 16) blocks = (int64_t)lc.sm_count * 16;
        k_check_csr<<<(unsigned)blocks, threads, 0, lc.stream>>>(X.ptr, X.idx, X.rows, X.cols, nnz, n, d_flags);
        SB_LAUNCH_CHECK(lc);
    }
    k_sorted_flag<<<1, 1, 0, lc.stream>>>(d_flags);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------
// Exclusive scan int32 -> OutT in three launches: tile sums, scan of <= 1024 tile sums, apply.
template <typename OutT>
__global__ void __launch_bounds__(256)
k_scan_tile_sums(const int32_t* __restrict__ in, int n, int tile, int64_t* __restrict__ sums) {
    __shared__ long long red[33];
    const int lo = blockIdx.x * tile, hi = min(n, lo + tile);
    long long s = 0;
    for (int t = lo + threadIdx.x; t < hi; t += blockDim.x) s += __ldg(in + t);
    long long total;
    block_excl_scan<long long>(s, red, &total);
    if (threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024)
k_scan_sums(int64_t* __restrict__ sums, int nb) {
    __shared__ long long red[33];
    long long v = (int)threadIdx.x < nb ? sums[threadIdx.x] : 0;
    long long total;
    long long ex = block_excl_scan<long long>(v, red, &total);
    if ((int)threadIdx.x < nb) sums[threadIdx.x] = ex;
    if (threadIdx.x == 0) sums[nb] = total;
}

template <typename OutT>
__global__ void __launch_bounds__(256)
k_scan_apply(const int32_t* __restrict__ in, OutT* __restrict__ out, int n, int tile, const int64_t* __restrict__ sums,
             int nb) {
    __shared__ long long red[33];
    const int lo = blockIdx.x * tile, hi = min(n, lo + tile);
    long long running = sums[blockIdx.x];
    for (int base = lo; base < hi; base += blockDim.x) {
        const int t = base + threadIdx.x;
        long long v = t < hi ? __ldg(in + t) : 0;
        long long total;
        long long ex = block_excl_scan<long long>(v, red, &total);
        if (t < hi) out[t] = (OutT)(running + ex);
        running += total;
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = (OutT)sums[nb];
}

// Small inputs: the whole scan in ONE block (each thread owns a contiguous chunk), one launch instead of three.
template <typename OutT>
__global__ void __launch_bounds__(1024)
k_scan_small(const int32_t* __restrict__ in, OutT* __restrict__ out, int n) {
    __shared__ long long red[33];
    const int per = (n + 1023) / 1024;
    const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
    long long s = 0;
    for (int t = lo; t < hi; ++t) s += __ldg(in + t);
    long long total;
    long long run = block_excl_scan<long long>(s, red, &total);
    for (int t = lo; t < hi; ++t) {
        out[t] = (OutT)run;
        run += __ldg(in + t);
    }
    if (threadIdx.x == 0) out[n] = (OutT)total;
}

template <typename OutT>
static cudaError_t launch_scan(const LaunchCtx& lc, const int32_t* in, OutT* out, int n, int64_t* d_tmp) {
    if (n <= 0) {
        return cudaMemsetAsync(out, 0, sizeof(OutT), lc.stream);
    }
    if (n <= 32768) {
        k_scan_small<OutT><<<1, 1024, 0, lc.stream>>>(in, out, n);
        SB_LAUNCH_CHECK(lc);
        return cudaSuccess;
    }
    int tile = (n + 1023) / 1024;
    tile = ((tile + 255) / 256) * 256;
    if (tile < 2048) tile = 2048;
    const int nb = (n + tile - 1) / tile;
    k_scan_tile_sums<OutT><<<nb, 256, 0, lc.stream>>>(in, n, tile, d_tmp);
    SB_LAUNCH_CHECK(lc);
    k_scan_sums<<<1, 1024, 0, lc.stream>>>(d_tmp, nb);
    SB_LAUNCH_CHECK(lc);
    k_scan_apply<OutT><<<nb, 256, 0, lc.stream>>>(in, out, n, tile, d_tmp, nb);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}
cudaError_t launch_scan_i64(const LaunchCtx& lc, const int32_t* in, int64_t* out, int n, int64_t* d_tmp) {
    return launch_scan<int64_t>(lc, in, out, n, d_tmp);
}
cudaError_t launch_scan_i32(const LaunchCtx& lc, const int32_t* in, int32_t* out, int n, int64_t* d_tmp) {
    return launch_scan<int32_t>(lc, in, out, n, d_tmp);
}

// ---------------------------------------------------------------------------------------------------
// Numeric binning by exact row nnz (one thread per row, block-aggregated list append).
__global__ void __launch_bounds__(256)
k_bin_by_nnz(const int32_t* __restrict__ nnz, int nrows, int32_t* __restrict__ lists,
             int32_t* __restrict__ cursor) {
    __shared__ int s_cnt[NUM_BINS], s_base[NUM_BINS];
    if (threadIdx.x < NUM_BINS) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    int bin = -1;
    if (r < nrows) {
        const int c = __ldg(nnz + r);
        if (c > 0) {
            if (c <= kWarpCap64) bin = NUM_W64;
            else if (c <= kWarpCap256) bin = NUM_W256;
            else if (c <= kWarpCap1K) bin = NUM_W1K;
            else bin = NUM_RANK;
        }
    }
    int local = 0;
    if (bin >= 0) local = atomicAdd(&s_cnt[bin], 1);
    __syncthreads();
    if (threadIdx.x < NUM_BINS) {
        const int c = s_cnt[threadIdx.x];
        s_base[threadIdx.x] = c ? atomicAdd(cursor + threadIdx.x, c) : 0;
    }
    __syncthreads();
    if (bin >= 0) lists[(size_t)bin * nrows + s_base[bin] + local] = r;
}

cudaError_t launch_bin_by_nnz(const LaunchCtx& lc, const int32_t* d_nnz, int nrows, int32_t* d_lists,
                              int32_t* d_cursor) {
    if (nrows <= 0) return cudaSuccess;
    k_bin_by_nnz<<<(nrows + 255) / 256, 256, 0, lc.stream>>>(d_nnz, nrows, d_lists, d_cursor);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------
// CSR transpose (used for H^T of the triple product): column histogram, scan (above), scatter.
__global__ void __launch_bounds__(256)
k_transpose_count(const int32_t* __restrict__ idx, int64_t nnz, int32_t* __restrict__ counts) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nnz) atomicAdd(counts + __ldg(idx + t), 1);
}
// LANES lanes per row of X (8 for short rows, 32 for rows of ~100+ entries); two entries per lane in flight
template <int LANES>
__global__ void __launch_bounds__(256)
k_transpose_fill(Csr X, const int32_t* __restrict__ t_ptr, int32_t* __restrict__ cursor, int32_t* __restrict__ t_idx,
                 double* __restrict__ t_val) {
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int gl = threadIdx.x % LANES;
    if (r >= X.rows) return;
    const int s = __ldg(X.ptr + r), e = __ldg(X.ptr + r + 1);
    for (int p = s + gl; p < e; p += 2 * LANES) {
        const int p2 = p + LANES;
        const int c0 = __ldg(X.idx + p);
        const int c1 = p2 < e ? __ldg(X.idx + p2) : -1;
        const double v0 = __ldg(X.val + p);
        const double v1 = p2 < e ? __ldg(X.val + p2) : 0.0;
        const int b0 = __ldg(t_ptr + c0);
        const int b1 = c1 >= 0 ? __ldg(t_ptr + c1) : 0;
        const int o0 = atomicAdd(cursor + c0, 1);
        const int o1 = c1 >= 0 ? atomicAdd(cursor + c1, 1) : 0;
        t_idx[b0 + o0] = r;
        t_val[b0 + o0] = v0;
        if (c1 >= 0) {
            t_idx[b1 + o1] = r;
            t_val[b1 + o1] = v1;
        }
    }
}
// ---------------------------------------------------------------------------------------------------
// Paneled transpose (triple product): rows [row_begin, row_end) of X are cut into panels of `panel_w` rows and the
// transposes of the panels are stored back to back as ONE CSR with np * X.cols rows -- row p * X.cols + c holds the
// entries (r, x_rc) of column c with r in panel p, each stored as ONE packed 32-bit word in t_pk next to its value:
// (r - first row of the panel) << 7 | (c & 127).  The low bits are all the banded-Q kernel needs of the column (its
// weight tables span at most 96 consecutive columns), so an entry costs 12 bytes of the stream instead of 16.
// t_ptr + p * X.cols is then the row-pointer array of panel p's transpose over the shared t_pk / t_val arrays.  The
// triple product accumulates one column panel of C at a time, so the slice of H^T it gathers from (a few tens of MB)
// stays L2 resident.
template <int LANES>
__global__ void __launch_bounds__(256)
k_transpose_count_panels(Csr X, int row_begin, int row_end, int panel_w, int32_t* __restrict__ counts) {
    const int r = row_begin + (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int gl = threadIdx.x % LANES;
    if (r >= row_end) return;
    int32_t* base = counts + (size_t)((r - row_begin) / panel_w) * X.cols;
    const int s = __ldg(X.ptr + r), e = __ldg(X.ptr + r + 1);
    for (int p = s + gl; p < e; p += LANES) atomicAdd(base + __ldg(X.idx + p), 1);
}

template <int LANES>
__global__ void __launch_bounds__(256)
k_transpose_fill_panels(Csr X, int row_begin, int row_end, int panel_w, const int32_t* __restrict__ t_ptr,
                        int32_t* __restrict__ cursor, uint32_t* __restrict__ t_pk, double* __restrict__ t_val) {
    const int r = row_begin + (blockIdx.x * blockDim.x + threadIdx.x) / LANES;
    const int gl = threadIdx.x % LANES;
    if (r >= row_end) return;
    const int panel = (r - row_begin) / panel_w;
    const size_t off = (size_t)panel * X.cols;
    const uint32_t rel = (uint32_t)(r - row_begin - panel * panel_w) << kPanelColBits;
    const int s = __ldg(X.ptr + r), e = __ldg(X.ptr + r + 1);
    for (int p = s + gl; p < e; p += 2 * LANES) {
        const int p2 = p + LANES;
        const int c0 = __ldg(X.idx + p);
        const int c1 = p2 < e ? __ldg(X.idx + p2) : -1;
        const double v0 = __ldg(X.val + p);
        const double v1 = p2 < e ? __ldg(X.val + p2) : 0.0;
        const int b0 = __ldg(t_ptr + off + c0);
        const int b1 = c1 >= 0 ? __ldg(t_ptr + off + c1) : 0;
        const int o0 = atomicAdd(cursor + off + c0, 1);
        const int o1 = c1 >= 0 ? atomicAdd(cursor + off + c1, 1) : 0;
        t_pk[b0 + o0] = rel | ((uint32_t)c0 & kPanelColMask);
        t_val[b0 + o0] = v0;
        if (c1 >= 0) {
            t_pk[b1 + o1] = rel | ((uint32_t)c1 & kPanelColMask);
            t_val[b1 + o1] = v1;
        }
    }
}

cudaError_t launch_transpose_count_panels(const LaunchCtx& lc, const Csr& X, int64_t nnz, int row_begin, int row_end,
                                          int panel_w, int32_t* d_counts) {
    const int rows = row_end - row_begin;
    if (rows <= 0) return cudaSuccess;
    if (nnz >= (int64_t)48 * X.rows)
        k_transpose_count_panels<32><<<(rows + 7) / 8, 256, 0, lc.stream>>>(X, row_begin, row_end, panel_w, d_counts);
    else
        k_transpose_count_panels<8><<<(rows + 31) / 32, 256, 0, lc.stream>>>(X, row_begin, row_end, panel_w, d_counts);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}
cudaError_t launch_transpose_fill_panels(const LaunchCtx& lc, const Csr& X, int64_t nnz, int row_begin, int row_end,
                                         int panel_w, const int32_t* t_ptr, int32_t* d_cursor, uint32_t* t_pk,
                                         double* t_val) {
    const int rows = row_end - row_begin;
    if (rows <= 0) return cudaSuccess;
    if (nnz >= (int64_t)48 * X.rows)
        k_transpose_fill_panels<32><<<(rows + 7) / 8, 256, 0, lc.stream>>>(X, row_begin, row_end, panel_w, t_ptr, d_cursor, t_pk, t_val);
    else
        k_transpose_fill_panels<8><<<(rows + 31) / 32, 256, 0, lc.stream>>>(X, row_begin, row_end, panel_w, t_ptr, d_cursor, t_pk, t_val);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

cudaError_t launch_transpose_count(const LaunchCtx& lc, const Csr& X, int64_t nnz, int32_t* d_counts) {
    if (nnz <= 0) return cudaSuccess;
    k_transpose_count<<<(unsigned)((nnz + 255) / 256), 256, 0, lc.stream>>>(X.idx, nnz, d_counts);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}
cudaError_t launch_transpose_fill(const LaunchCtx& lc, const Csr& X, int64_t nnz, const int32_t* t_ptr,
                                  int32_t* d_cursor, int32_t* t_idx, double* t_val) {
    if (X.rows <= 0) return cudaSuccess;
    if (nnz >= (int64_t)48 * X.rows)
        k_transpose_fill<32><<<(X.rows + 7) / 8, 256, 0, lc.stream>>>(X, t_ptr, d_cursor, t_idx, t_val);
    else
        k_transpose_fill<8><<<(X.rows + 31) / 32, 256, 0, lc.stream>>>(X, t_ptr, d_cursor, t_idx, t_val);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

// Sort the entries of every row by column, in place, ascending or descending.  DESCENDING is used on H^T: the
// triple product's upper-triangle contraction then reads a row of H^T from the front and stops at the first entry
// below the diagonal.  ASCENDING is the device-side canonicalisation of an unsorted right operand (SURVEY.md 8(f).2;
// the reference coerces on the host, sparse_matrix_mult/matrix_ops.py:307-310): the column-window kernels can then
// binary-search instead of filtering every entry.  Duplicates stay (they are summed by the accumulators).
//   k_sort_rows            one thread per row, insertion sort, rows of <= kSortShort entries (the common case: a row
//                          of H^T holds the entries of one column of H); longer rows are appended to long_list
//                          (long_list[0] = count, row ids from long_list[1])
//   k_sort_long_rows       one block per listed row: bitonic network with every compare-exchange in the same
//                          direction (first step of each merge mirrors the block), which sorts ANY length with
//                          the out-of-range partners skipped; runs in global memory (rows of any length)
constexpr int kSortShort = 32;

template <bool DESC>
__global__ void __launch_bounds__(256)
k_sort_rows(int rows, const int32_t* __restrict__ ptr, int32_t* __restrict__ idx, double* __restrict__ val,
            int32_t* __restrict__ long_list) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int s = ptr[r], e = ptr[r + 1];
    if (e - s <= 1) return;
    if (e - s > kSortShort) { long_list[1 + atomicAdd(long_list, 1)] = r; return; }
    for (int a = s + 1; a < e; ++a) {
        const int k = idx[a];
        const double v = val[a];
        int b = a - 1;
        while (b >= s && (DESC ? idx[b] < k : idx[b] > k)) { idx[b + 1] = idx[b]; val[b + 1] = val[b]; --b; }
        idx[b + 1] = k;
        val[b + 1] = v;
    }
}

template <bool DESC>
__global__ void __launch_bounds__(256)
k_sort_long_rows(const int32_t* __restrict__ ptr, int32_t* idx, double* val, const int32_t* __restrict__ long_list) {
    const int count = long_list[0];
    for (int item = blockIdx.x; item < count; item += gridDim.x) {
        const int r = long_list[1 + item];
        const int s = ptr[r], n = ptr[r + 1] - s;
        int32_t* ki = idx + s;
        double* vi = val + s;
        int np2 = 1;
        while (np2 < n) np2 <<= 1;
        const int half = np2 >> 1;
        auto cmpx = [&](int i, int l) {              // DESC: keep the larger column at the lower position
            const int a = ki[i], b = ki[l];
            if (DESC ? a < b : a > b) {
                ki[i] = b; ki[l] = a;
                const double va = vi[i], vb = vi[l];
                vi[i] = vb; vi[l] = va;
            }
        };
        for (int k = 2; k <= np2; k <<= 1) {
            const int hk = k >> 1;
            for (int t = threadIdx.x; t < half; t += blockDim.x) {       // mirror step: i <-> block_end - offset
                const int blk = t / hk, off = t % hk;
                const int i = blk * k + off, l = blk * k + (k - 1 - off);
                if (l < n) cmpx(i, l);
            }
            __syncthreads();
            for (int j = k >> 2; j > 0; j >>= 1) {
                for (int t = threadIdx.x; t < half; t += blockDim.x) {
                    const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                    const int l = i | j;
                    if (l < n) cmpx(i, l);
                }
                __syncthreads();
            }
        }
    }
}

cudaError_t launch_sort_rows(const LaunchCtx& lc, int rows, const int32_t* ptr, int32_t* idx, double* val,
                             int32_t* d_long_list, bool descending) {
    if (rows <= 0) return cudaSuccess;
    cudaError_t e = cudaMemsetAsync(d_long_list, 0, sizeof(int32_t), lc.stream);
    if (e != cudaSuccess) return e;
    if (descending) k_sort_rows<true><<<(rows + 255) / 256, 256, 0, lc.stream>>>(rows, ptr, idx, val, d_long_list);
    else k_sort_rows<false><<<(rows + 255) / 256, 256, 0, lc.stream>>>(rows, ptr, idx, val, d_long_list);
    SB_LAUNCH_CHECK(lc);
    if (descending) k_sort_long_rows<true><<<lc.sm_count * 4, 256, 0, lc.stream>>>(ptr, idx, val, d_long_list);
    else k_sort_long_rows<false><<<lc.sm_count * 4, 256, 0, lc.stream>>>(ptr, idx, val, d_long_list);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_narrow(const int64_t* __restrict__ in, int32_t* __restrict__ out, int n) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) out[t] = (int32_t)in[t];
}
cudaError_t launch_narrow_indptr(const LaunchCtx& lc, const int64_t* in, int32_t* out, int n) {
    if (n <= 0) return cudaSuccess;
    k_narrow<<<(n + 255) / 256, 256, 0, lc.stream>>>(in, out, n);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------
// costs[i] += c: the part of a row's cost that does not depend on its products (a dense output row is written in
// full, zeros included: for BASELINE config 2 that is all of the time).
__global__ void __launch_bounds__(256)
k_add_const(int64_t* __restrict__ costs, int n, long long c) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) costs[t] += c;
}
cudaError_t launch_add_const(const LaunchCtx& lc, int64_t* d_costs, int n, long long c) {
    if (n <= 0 || c == 0) return cudaSuccess;
    k_add_const<<<(n + 255) / 256, 256, 0, lc.stream>>>(d_costs, n, c);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------
// Cost of row i of the triple product H Q H^T for the flop-balanced multi-GPU split:
//   P1_i = sum over (i,j) in H of nnz(Q[j,:])                         (expansion of H[i,:] Q)
//   P2_i = sum over those (j,c) of nnz(H^T[c,:])                      (entries of H^T the contraction streams)
// cost_i = a * P1_i * panels(i) + b * P2_i * (n - i) / n + c * n: the expansion is redone for every column panel
// the row takes part in (a; upper mode: the panels right of the diagonal), the entries of H^T in those panels --
// the fraction (n - i) / n of all of them in upper mode -- are multiplied and accumulated (b), and every row of C is
// written in full, zeros included (c).  (b, c) were fitted jointly to the 15 per-rank step times of cfg 5 on 1, 2, 4
// and 8 B200s (profiles/r2/SUMMARY.md): 0.454 and 0.289 ms per 10^8 units with the banded-Q kernel, i.e. c / b = 0.64,
// together with the transposition term the partition adds per block (kTripleTailCoeff, ctx.h); the fit leaves +-3 %,
// the spread between the GPUs of one box.  a is not identifiable on that matrix (P1_i is almost the same for every
// row, so it is collinear with c) and stays small; the general kernel pays ~100 instructions per 32 products (a = 2, a
// guess).  SPGEMM_B200_TRIPLE_COST="a,b,c" overrides them.  One warp per row.
__global__ void __launch_bounds__(256)
k_triple_costs(Csr H, Csr Q, Csr Ht, int upper_only, int np, int panel_w, double ca, double cb, double cc,
               int64_t* __restrict__ costs) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= H.rows) return;
    long long p1 = 0, p2 = 0;
    expand_row_warp<false>(H, Q, __ldg(H.ptr + row), __ldg(H.ptr + row + 1), 0, 0, false, false,
                           [&](int c, double) {
                               ++p1;
                               p2 += __ldg(Ht.ptr + c + 1) - __ldg(Ht.ptr + c);
                           });
    p1 = warp_sum(p1);
    p2 = warp_sum(p2);
    if (lane_id() == 0) {
        const double keep = upper_only ? (double)(H.rows - row) / (double)H.rows : 1.0;
        const int panels = upper_only ? np - row / panel_w : np;
        costs[row] = (long long)(ca * (double)p1 * (double)(panels > 1 ? panels : 1) + cb * (double)p2 * keep +
                                 cc * (double)H.rows);
    }
}
cudaError_t launch_triple_costs(const LaunchCtx& lc, const Csr& H, const Csr& Q, const Csr& Ht, bool upper_only,
                                bool q_runs, int np, int panel_w, int64_t* d_costs) {
    if (H.rows <= 0) return cudaSuccess;
    double ca = q_runs ? 0.05 : 2.0, cb = 1.0, cc = 0.64;
    if (const char* v = getenv("SPGEMM_B200_TRIPLE_COST")) sscanf(v, "%lf,%lf,%lf", &ca, &cb, &cc);
    k_triple_costs<<<(H.rows + 7) / 8, 256, 0, lc.stream>>>(H, Q, Ht, upper_only ? 1 : 0, np, panel_w > 0 ? panel_w : 1,
                                                           ca, cb, cc, d_costs);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

}  // namespace sb
