// spgemm_sparse.cu -- sparse-output SpGEMM: symbolic (row nnz of C) and numeric (Gustavson) phases.
//
// Replaces the reference's per-thread Gustavson with a dense position-marker array
// (/root/reference/src/sparsework.cpp:56-129 and :201-280) and its serial stitch
// (/root/reference/src/sparse_sparse_sparse.cpp:265-291).  Here every row of C is owned by one warp or one
// thread block, chosen by the row's cost bin:
//   warp bins : open-addressing hash table in shared memory (64 / 256 / 1024 slots per warp), compacted and
//               bitonic-sorted by column inside the warp
//   block bin : occupancy bitmap of the row in shared memory; a popcount prefix over it gives every column its
//               rank in the sorted row, so values are accumulated in place in C (no hash table, no sort)
// Output rows are written at their final position (int64 offsets from the scan of the symbolic counts):
// no stitch pass.  Entries whose value cancels to zero stay (they are structural in the reference too).
#include <cstdlib>

#include "internal.h"

namespace sb {

#define SB_LAUNCH_CHECK(lc)                  \
    do {                                     \
        ++*(lc).launches;                    \
        cudaError_t e_ = cudaGetLastError(); \
        if (e_ != cudaSuccess) return e_;    \
    } while (0)

constexpr int kKeyMax = 0x7fffffff;

// ---------------------------------------------------------------------------------------------------
// hash-table primitives (shared memory)
__device__ __forceinline__ bool hash_insert_key(int* keys, unsigned size, int c) {
    unsigned h = hash_slot(c, size);
    while (true) {
        const int k = *((volatile int*)(keys + h));
        if (k == c) return false;
        if (k == kEmpty) {
            const int old = atomicCAS(keys + h, kEmpty, c);
            if (old == kEmpty) return true;
            if (old == c) return false;
        }
        h = (h + 1 == size) ? 0 : h + 1;
    }
}

__device__ __forceinline__ void hash_accumulate(int* keys, double* vals, unsigned size, int c, double v) {
    unsigned h = hash_slot(c, size);
    while (true) {
        const int k = *((volatile int*)(keys + h));
        if (k == c) break;
        if (k == kEmpty) {
            const int old = atomicCAS(keys + h, kEmpty, c);
            if (old == kEmpty || old == c) break;
        }
        h = (h + 1 == size) ? 0 : h + 1;
    }
    atomicAdd(vals + h, v);
}

// Bitonic sort of n (power of two) (key, value) pairs in shared memory by `nthreads` cooperating threads.
template <bool BLOCK>
__device__ __forceinline__ void group_sync() {
    if (BLOCK) __syncthreads(); else __syncwarp();
}

template <bool BLOCK>
__device__ __forceinline__ void bitonic_sort_pairs(int* keys, double* vals, int n, int tid, int nthreads) {
    for (int k = 2; k <= n; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = tid; t < (n >> 1); t += nthreads) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool up = (i & k) == 0;
                const int ka = keys[i], kb = keys[l];
                if ((ka > kb) == up) {
                    keys[i] = kb; keys[l] = ka;
                    const double va = vals[i], vb = vals[l];
                    vals[i] = vb; vals[l] = va;
                }
            }
            group_sync<BLOCK>();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// Symbolic, warp per row.
template <int SLOTS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_symbolic_warp(Csr A, Csr B, int row_begin, int upper_only, const int32_t* __restrict__ b_sorted_flag,
                const int32_t* __restrict__ list, int count, int32_t* __restrict__ nnz) {
    __shared__ int s_keys[WARPS * SLOTS];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int* keys = s_keys + warp * SLOTS;
    const bool b_sorted = *b_sorted_flag != 0;
    for (int w = blockIdx.x * WARPS + warp; w < count; w += gridDim.x * WARPS) {
        const int r = __ldg(list + w), i = row_begin + r;
        for (int t = lane; t < SLOTS; t += 32) keys[t] = kEmpty;
        __syncwarp();
        int found = 0;
        expand_row_warp<false>(A, B, __ldg(A.ptr + i), __ldg(A.ptr + i + 1), upper_only ? i : 0, B.cols,
                               upper_only != 0, b_sorted,
                               [&](int c, double) { found += hash_insert_key(keys, SLOTS, c) ? 1 : 0; });
        found = warp_sum(found);
        if (lane == 0) nnz[r] = found;
        __syncwarp();
    }
}

// Symbolic, block per row, occupancy bitmap over column windows of `window_bits` columns.
__global__ void __launch_bounds__(512)
k_symbolic_bitmap(Csr A, Csr B, int row_begin, int upper_only, const int32_t* __restrict__ b_sorted_flag,
                  const int32_t* __restrict__ list, int count, int window_bits, int32_t* __restrict__ nnz,
                  int32_t* __restrict__ work_counter) {
    extern __shared__ unsigned s_bits[];
    __shared__ int s_item;
    __shared__ int s_red[33];
    __shared__ SegScratch<512> s_seg;
    const bool b_sorted = *b_sorted_flag != 0;
    const int n = B.cols;
    while (true) {
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= count) break;
        const int r = __ldg(list + item), i = row_begin + r;
        const int a_begin = __ldg(A.ptr + i), a_end = __ldg(A.ptr + i + 1);
        const int lo = upper_only ? i : 0;
        int total = 0;
        for (int w0 = (lo / window_bits) * window_bits; w0 < n; w0 += window_bits) {
            const int wl = max(w0, lo), wh = min(w0 + window_bits, n);
            const int words = (wh - w0 + 31) >> 5;
            for (int t = threadIdx.x; t < words; t += blockDim.x) s_bits[t] = 0u;
            __syncthreads();
            const bool windowed = upper_only || window_bits < n;
            expand_row_block<false>(A, B, a_begin, a_end, wl, wh, windowed, b_sorted, s_seg, [&](int c, double) {
                const int o = c - w0;
                const unsigned m = 1u << (o & 31);
                if (!(*((volatile unsigned*)(s_bits + (o >> 5))) & m)) atomicOr(s_bits + (o >> 5), m);
            });
            __syncthreads();
            int cnt = 0;
            for (int t = threadIdx.x; t < words; t += blockDim.x) cnt += __popc(s_bits[t]);
            int sum;
            block_excl_scan<int>(cnt, s_red, &sum);
            total += sum;
        }
        if (threadIdx.x == 0) nnz[r] = total;
    }
}

// ---------------------------------------------------------------------------------------------------
// Numeric, warp per row: hash accumulate, compact in place, bitonic sort, coalesced write-out.
template <int SLOTS, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
k_numeric_warp(Csr A, Csr B, int row_begin, int upper_only, const int32_t* __restrict__ b_sorted_flag,
               const int32_t* __restrict__ list, int count, const int64_t* __restrict__ c_ptr,
               int32_t* __restrict__ c_idx, double* __restrict__ c_val) {
    __shared__ double s_vals[WARPS * SLOTS];
    __shared__ int s_keys[WARPS * SLOTS];
    const int lane = lane_id(), warp = threadIdx.x >> 5;
    int* keys = s_keys + warp * SLOTS;
    double* vals = s_vals + warp * SLOTS;
    const bool b_sorted = *b_sorted_flag != 0;
    for (int w = blockIdx.x * WARPS + warp; w < count; w += gridDim.x * WARPS) {
        const int r = __ldg(list + w), i = row_begin + r;
        for (int t = lane; t < SLOTS; t += 32) { keys[t] = kEmpty; vals[t] = 0.0; }
        __syncwarp();
        expand_row_warp<true>(A, B, __ldg(A.ptr + i), __ldg(A.ptr + i + 1), upper_only ? i : 0, B.cols,
                              upper_only != 0, b_sorted,
                              [&](int c, double v) { hash_accumulate(keys, vals, SLOTS, c, v); });
        __syncwarp();
        // compact occupied slots to the front (stable in slot order, in place)
        int fill = 0;
        for (int base = 0; base < SLOTS; base += 32) {
            const int k = keys[base + lane];
            const double v = vals[base + lane];
            const unsigned m = __ballot_sync(FULL, k != kEmpty);
            __syncwarp();
            if (k != kEmpty) {
                const int pos = fill + __popc(m & ((1u << lane) - 1u));
                keys[pos] = k;
                vals[pos] = v;
            }
            fill += __popc(m);
            __syncwarp();
        }
        int n2 = 2;
        while (n2 < fill) n2 <<= 1;
        for (int t = fill + lane; t < n2; t += 32) keys[t] = kKeyMax;
        __syncwarp();
        bitonic_sort_pairs<false>(keys, vals, n2, lane, 32);
        const int64_t off = __ldg(c_ptr + r);
        for (int t = lane; t < fill; t += 32) {
            c_idx[off + t] = keys[t];
            c_val[off + t] = vals[t];
        }
        __syncwarp();
    }
}

// Numeric, block per row, for every row beyond the warp bins ("rank" kernel).  No hash table and no sort:
//   pass 1  marks the row's columns in an occupancy bitmap in shared memory (as the symbolic phase did);
//   prefix  exclusive popcount prefix per bitmap word: rank(c) = number of occupied columns < c, i.e. the
//           position of column c in the sorted output row ({bits, prefix} pairs: one 64-bit shared load);
//   emit    the sorted column indices are written straight from the bitmap;
//   pass 2  the row's values are accumulated BY RANK in a compact shared-memory array of `cap` doubles (shared
//           float64 atomics sustain ~2.7x the rate of L2 reductions on B200, scripts/micro/atomic_bw.cu) and
//           leave with coalesced stores.  Rows with more than `cap` entries take several rank windows; a rank
//           window is a column range, so with sorted B each window reads only its own part of the rows of B.
// Shared memory per block: table[W/32] of {bits, prefix} (W = column window, all columns when they fit) and
// vals[cap].  Wider matrices take several column windows.
template <bool SMEM_ACC>
__global__ void __launch_bounds__(512)
k_numeric_rank(Csr A, Csr B, int row_begin, int upper_only, const int32_t* __restrict__ b_sorted_flag,
               const int32_t* __restrict__ list, int count, int window, int cap, const int64_t* __restrict__ c_ptr,
               int32_t* __restrict__ c_idx, double* __restrict__ c_val, int32_t* __restrict__ work_counter) {
    extern __shared__ double s_dynd[];
    double* vals = s_dynd;
    uint2* table = reinterpret_cast<uint2*>(s_dynd + cap);       // .x = occupancy bits, .y = exclusive prefix
    unsigned* table_u = reinterpret_cast<unsigned*>(table);
    __shared__ int s_item;
    __shared__ int s_red[33];
    __shared__ int s_bound[2];
    __shared__ SegScratch<512> s_seg;
    const bool b_sorted = *b_sorted_flag != 0;
    const int n = B.cols;
    while (true) {
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= count) break;
        const int r = __ldg(list + item), i = row_begin + r;
        const int a_begin = __ldg(A.ptr + i), a_end = __ldg(A.ptr + i + 1);
        const int lo = upper_only ? i : 0;
        int64_t out = __ldg(c_ptr + r);
        for (int w0 = (lo / window) * window; w0 < n; w0 += window) {
            const int wl = max(w0, lo), wh = min(w0 + window, n);
            const int words = (wh - w0 + 31) >> 5;
            const bool col_windowed = upper_only || window < n;
            for (int t = threadIdx.x; t < words; t += blockDim.x) table[t] = make_uint2(0u, 0u);
            __syncthreads();
            // pass 1: occupancy
            expand_row_block<false>(A, B, a_begin, a_end, wl, wh, col_windowed, b_sorted, s_seg, [&](int c, double) {
                const int o = c - w0;
                const unsigned m = 1u << (o & 31);
                unsigned* wp = table_u + 2 * (o >> 5);
                if (!(*((volatile unsigned*)wp) & m)) atomicOr(wp, m);
            });
            __syncthreads();
            // exclusive popcount prefix over the words
            int nnz_w = 0;
            for (int base = 0; base < words; base += blockDim.x) {
                const int w = base + threadIdx.x;
                const int pc = w < words ? __popc(table[w].x) : 0;
                int tot;
                const int ex = block_excl_scan<int>(pc, s_red, &tot);
                if (w < words) table[w].y = (unsigned)(nnz_w + ex);
                nnz_w += tot;
            }
            __syncthreads();
            // emit the sorted column indices of this column window
            for (int w = threadIdx.x; w < words; w += blockDim.x) {
                const uint2 e = table[w];
                unsigned word = e.x;
                int64_t pos = out + e.y;
                while (word) {
                    const int b = __ffs(word) - 1;
                    word &= word - 1;
                    c_idx[pos++] = w0 + (w << 5) + b;
                }
            }
            if (!SMEM_ACC) {
                // pass 2, variant: accumulate straight into C.val with L2 reductions (no rank windows)
                double* gv = c_val + out;
                for (int t = threadIdx.x; t < nnz_w; t += blockDim.x) gv[t] = 0.0;
                __syncthreads();
                expand_row_block<true>(A, B, a_begin, a_end, wl, wh, col_windowed, b_sorted, s_seg,
                                       [&](int c, double v) {
                                           const int o = c - w0;
                                           const uint2 e = table[o >> 5];
                                           atomicAdd(gv + (int)e.y + __popc(e.x & ((1u << (o & 31)) - 1u)), v);
                                       });
                out += nnz_w;
                __syncthreads();
                continue;
            }
            // pass 2: values, one rank window of <= cap entries at a time
            const int nwin = (nnz_w + cap - 1) / cap;
            for (int k = 0; k < nwin; ++k) {
                const int r0 = k * cap, r1 = min(nnz_w, r0 + cap);
                if (nwin > 1) {
                    // column bounds of the rank window: column of rank r0 / r1 (threads 0 and 32 search)
                    if (threadIdx.x == 0 || threadIdx.x == 32) {
                        const int target = threadIdx.x == 0 ? r0 : r1;
                        int col;
                        if (target >= nnz_w) col = wh;
                        else {
                            int a = 0, b = words;            // largest w with prefix[w] <= target
                            while (b - a > 1) {
                                const int mid = (a + b) >> 1;
                                if ((int)table[mid].y <= target) a = mid; else b = mid;
                            }
                            unsigned word = table[a].x;
                            for (int skip = target - (int)table[a].y; skip > 0; --skip) word &= word - 1;
                            col = w0 + (a << 5) + __ffs(word) - 1;
                        }
                        s_bound[threadIdx.x == 0 ? 0 : 1] = col;
                    }
                    __syncthreads();
                }
                const int cl = nwin > 1 ? max(s_bound[0], wl) : wl;
                const int ch = nwin > 1 ? s_bound[1] : wh;
                for (int t = threadIdx.x; t < r1 - r0; t += blockDim.x) vals[t] = 0.0;
                __syncthreads();
                expand_row_block<true>(A, B, a_begin, a_end, cl, ch, col_windowed || nwin > 1, b_sorted, s_seg,
                                       [&](int c, double v) {
                                           const int o = c - w0;
                                           const uint2 e = table[o >> 5];
                                           const int rank = (int)e.y + __popc(e.x & ((1u << (o & 31)) - 1u)) - r0;
                                           atomicAdd(vals + rank, v);
                                       },
                                       NoHook(), nwin > 2 ? 4 : kClipMin);
                __syncthreads();
                double* dst = c_val + out + r0;
                for (int t = threadIdx.x; t < r1 - r0; t += blockDim.x) dst[t] = vals[t];
                __syncthreads();
            }
            out += nnz_w;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
static size_t g_smem_optin = 0;

cudaError_t sparse_kernels_configure() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    int optin = 0;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    g_smem_optin = (size_t)optin;
    e = cudaFuncSetAttribute(k_symbolic_bitmap, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_numeric_rank<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_numeric_rank<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
    return e;
}

static inline int grid_for(int items, int per_block, int cap) {
    int g = (items + per_block - 1) / per_block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return g;
}

cudaError_t launch_symbolic(const LaunchCtx& lc, const SparseJob& job, const int32_t* d_lists, const int32_t* h_counts,
                            int32_t* d_nnz, int32_t* d_work_counter) {
    const int up = job.upper_only ? 1 : 0;
    const size_t stride = (size_t)job.nrows;
    const int cap = lc.sm_count * 16;
    if (h_counts[SYM_W64]) {
        constexpr int W = 8;
        k_symbolic_warp<64, W><<<grid_for(h_counts[SYM_W64], W, cap), W * 32, 0, lc.stream>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_W64 * stride, h_counts[SYM_W64], d_nnz);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[SYM_W256]) {
        constexpr int W = 8;
        k_symbolic_warp<256, W><<<grid_for(h_counts[SYM_W256], W, cap), W * 32, 0, lc.stream>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_W256 * stride, h_counts[SYM_W256], d_nnz);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[SYM_W1K]) {
        constexpr int W = 8;
        k_symbolic_warp<1024, W><<<grid_for(h_counts[SYM_W1K], W, cap), W * 32, 0, lc.stream>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_W1K * stride, h_counts[SYM_W1K], d_nnz);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[SYM_BITMAP]) {
        // window = all columns when they fit the per-block shared memory, else the largest multiple of 32 bits
        const size_t max_bits = (g_smem_optin - 20480) * 8;
        size_t window_bits = (size_t)job.B.cols;
        if (window_bits > max_bits) window_bits = max_bits & ~(size_t)1023;
        if (window_bits < 32) window_bits = 32;
        window_bits = (window_bits + 31) & ~(size_t)31;
        const size_t smem = window_bits / 8;
        // several blocks per SM when the bitmap is small
        int per_sm = (int)((g_smem_optin) / (smem + 10240));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        cudaError_t e = cudaMemsetAsync(d_work_counter, 0, sizeof(int32_t), lc.stream);
        if (e != cudaSuccess) return e;
        k_symbolic_bitmap<<<grid_for(h_counts[SYM_BITMAP], 1, lc.sm_count * per_sm), 512, smem, lc.stream>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_BITMAP * stride, h_counts[SYM_BITMAP],
            (int)window_bits, d_nnz, d_work_counter);
        SB_LAUNCH_CHECK(lc);
    }
    return cudaSuccess;
}

cudaError_t launch_numeric(const LaunchCtx& lc, const SparseJob& job, const int32_t* d_lists, const int32_t* h_counts,
                           const int64_t* c_ptr, int32_t* c_idx, double* c_val, int32_t* d_work_counter) {
    const int up = job.upper_only ? 1 : 0;
    const size_t stride = (size_t)job.nrows;
    const int cap = lc.sm_count * 16;
    if (h_counts[NUM_W64]) {
        constexpr int W = 8;
        k_numeric_warp<64, W><<<grid_for(h_counts[NUM_W64], W, cap), W * 32, 0, lc.stream>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_W64 * stride, h_counts[NUM_W64], c_ptr,
            c_idx, c_val);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[NUM_W256]) {
        constexpr int W = 8;
        k_numeric_warp<256, W><<<grid_for(h_counts[NUM_W256], W, cap), W * 32, 0, lc.stream>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_W256 * stride, h_counts[NUM_W256], c_ptr,
            c_idx, c_val);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[NUM_W1K]) {
        constexpr int W = 4;
        k_numeric_warp<1024, W><<<grid_for(h_counts[NUM_W1K], W, cap), W * 32, 0, lc.stream>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_W1K * stride, h_counts[NUM_W1K], c_ptr,
            c_idx, c_val);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[NUM_RANK]) {
        cudaError_t e = cudaMemsetAsync(d_work_counter, 0, sizeof(int32_t), lc.stream);
        if (e != cudaSuccess) return e;
        // column window: all columns when the {bits, prefix} table (8 B per 32 columns) leaves room for a useful
        // value array, else 2^19 columns (128 KB table)
        const size_t max_dyn = g_smem_optin - 20480;                     // what sparse_kernels_configure allows
        const size_t two_per_sm = g_smem_optin / 2 - 11264;              // dynamic bytes that still fit 2 blocks/SM
        int64_t window = ((int64_t)job.B.cols + 31) & ~(int64_t)31;
        if ((size_t)(window / 4) + 4096 * 8 > max_dyn) window = 1 << 19;
        const size_t table_bytes = (size_t)(window / 4);
        const size_t avail = (table_bytes + 4096 * 8 <= two_per_sm ? two_per_sm : max_dyn) - table_bytes;
        int cap = (int)(avail / 8);
        if (cap > 12288) cap = 12288;
        const char* env_cap = getenv("SPGEMM_B200_RANK_CAP");
        if (env_cap && atoi(env_cap) > 0 && atoi(env_cap) < cap) cap = atoi(env_cap);
        cap &= ~31;
        const char* env_mode = getenv("SPGEMM_B200_RANK_MODE");
        const bool smem_acc = env_mode && atoi(env_mode) == 1;      // default: L2 reductions (faster as measured)
        if (!smem_acc) cap = 0;
        const size_t smem = (size_t)cap * 8 + table_bytes;
        int per_sm = (int)((g_smem_optin + 1024) / (smem + 10240));
        if (per_sm > 4) per_sm = 4;
        if (per_sm < 1) per_sm = 1;
        if (smem_acc)
            k_numeric_rank<true><<<grid_for(h_counts[NUM_RANK], 1, lc.sm_count * per_sm), 512, smem, lc.stream>>>(
                job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_RANK * stride, h_counts[NUM_RANK],
                (int)window, cap, c_ptr, c_idx, c_val, d_work_counter);
        else
            k_numeric_rank<false><<<grid_for(h_counts[NUM_RANK], 1, lc.sm_count * per_sm), 512, smem, lc.stream>>>(
                job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_RANK * stride, h_counts[NUM_RANK],
                (int)window, cap, c_ptr, c_idx, c_val, d_work_counter);
        SB_LAUNCH_CHECK(lc);
    }
    return cudaSuccess;
}

}  // namespace sb
