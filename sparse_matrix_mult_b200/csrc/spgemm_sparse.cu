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
// When the whole row fits one window and `saved` is given, the finished bitmap is also stored to global memory
// (saved.words words per slot, slots handed out by an atomic counter until they run out; saved.slot_of_row[r] = the
// slot or -1): the numeric rank kernel then loads it instead of rebuilding it with a second pass over the products.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
k_symbolic_bitmap(Csr A, Csr B, int row_begin, int upper_only, const int32_t* __restrict__ b_sorted_flag,
                  const int32_t* __restrict__ list, int count, int window_bits, int32_t* __restrict__ nnz,
                  int32_t* __restrict__ work_counter, SavedBitmaps saved) {
    extern __shared__ unsigned s_bits[];
    __shared__ int s_item;
    __shared__ int s_slot;
    __shared__ int s_red[33];
    __shared__ SegScratch<THREADS> s_seg;
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
            if (saved.bits != nullptr && window_bits >= n) {       // single window: keep the bitmap for the numeric phase
                if (threadIdx.x == 0) {
                    const int slot = sum > kWarpCap1K ? atomicAdd(saved.counter, 1) : saved.slots;   // warp-bin rows: no
                    s_slot = slot < saved.slots ? slot : -1;
                    saved.slot_of_row[r] = s_slot;
                }
                __syncthreads();
                if (s_slot >= 0) {
                    unsigned* dst = saved.bits + (size_t)s_slot * saved.words;
                    for (int t = threadIdx.x; t < saved.words; t += blockDim.x) dst[t] = t < words ? s_bits[t] : 0u;
                }
                __syncthreads();
            }
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
//   prefix  exclusive popcount prefix over the bitmap: rank(c) = number of occupied columns < c, i.e. the
//           position of column c in the sorted output row;
//   emit    the sorted column indices are written straight from the bitmap, the row's values are zeroed;
//   pass 2  every product is added into C.val[row_offset + rank(col)] with a float64 reduction that resolves in
//           L2 (native RED.ADD.F64; the row's slice of C.val was just written, so it is L2 resident).
// Two table layouts:
//   COMPACT = false  {bits, prefix} pair per 32 columns (8 B): one 64-bit shared load per rank.  65,536 columns
//                    need 16 KB, so several blocks share an SM.
//   COMPACT = true   bits[W/32] plus one prefix per FOUR words (5 B per 32 columns): 1,048,576 columns fit one
//                    160 KB table, one 1024-thread block per SM, a 128-bit shared load per rank.
// Matrices wider than the table take several column windows.
// (Two variants measured slower and were removed -- DESIGN.md section 7: accumulating by rank in shared memory, in
//  rank windows (round 1), and a block-wide shared-memory hash table of (column, value) scattered by bitmap rank,
//  one traversal of the products instead of two (round 2: cfg 4r numeric 6.7 ms against 3.6 ms, barrier stalls and
//  compare-and-swap contention on the hub columns of power-law inputs; profiles/r2/SUMMARY.md).)
template <bool COMPACT>
struct RankTable {
    unsigned* bits;      // COMPACT: bits[words];            else: interleaved {bits, prefix}
    unsigned* pre;       // COMPACT: prefix per 4 words;     else: unused
    __device__ __forceinline__ unsigned* word_ptr(int w) const { return COMPACT ? bits + w : bits + 2 * w; }
    __device__ __forceinline__ int rank(int o) const {
        const int w = o >> 5;
        const unsigned below = (1u << (o & 31)) - 1u;
        if (!COMPACT) {
            const uint2 e = reinterpret_cast<const uint2*>(bits)[w];
            return (int)e.y + __popc(e.x & below);
        }
        const uint4 b = reinterpret_cast<const uint4*>(bits)[w >> 2];
        const int k = w & 3;
        const unsigned cur = k == 0 ? b.x : k == 1 ? b.y : k == 2 ? b.z : b.w;
        int r = (int)pre[w >> 2] + __popc(cur & below);
        if (k > 0) r += __popc(b.x);
        if (k > 1) r += __popc(b.y);
        if (k > 2) r += __popc(b.z);
        return r;
    }
};

template <bool COMPACT, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_numeric_rank(Csr A, Csr B, int row_begin, int upper_only, const int32_t* __restrict__ b_sorted_flag,
               const int32_t* __restrict__ list, int count, int stride, int window,
               const int64_t* __restrict__ c_ptr, int32_t* __restrict__ c_idx, double* __restrict__ c_val,
               int32_t* __restrict__ work_counter, SavedBitmaps saved) {
    extern __shared__ unsigned s_dynu[];
    RankTable<COMPACT> tab;
    tab.bits = s_dynu;
    tab.pre = s_dynu + (window >> 5);            // COMPACT only
    __shared__ int s_item;
    __shared__ int s_red[33];
    __shared__ SegScratch<THREADS> s_seg;
    const bool b_sorted = *b_sorted_flag != 0;
    const int n = B.cols;
    while (true) {
        if (threadIdx.x == 0) s_item = atomicAdd(work_counter, 1);
        __syncthreads();
        const int item = s_item;
        __syncthreads();
        if (item >= count) break;
        // Tickets are mapped to list positions by a stride coprime with `count`: the list is roughly in row order
        // and on power-law inputs the heavy rows sit together; scattered, the rows in flight are a random sample
        // and fewer multi-megabyte value slices compete for L2 at once (cfg 4: 257 -> 244 ms).
        const int r = __ldg(list + (int)(((long long)item * stride) % count)), i = row_begin + r;
        const int a_begin = __ldg(A.ptr + i), a_end = __ldg(A.ptr + i + 1);
        const int lo = upper_only ? i : 0;
        int64_t out = __ldg(c_ptr + r);
        // the symbolic phase may have kept this row's bitmap (single-window rows only)
        const int slot = (saved.bits != nullptr && window >= n) ? __ldg(saved.slot_of_row + r) : -1;
        for (int w0 = (lo / window) * window; w0 < n; w0 += window) {
            const int wl = max(w0, lo), wh = min(w0 + window, n);
            const int words = (((wh - w0 + 31) >> 5) + 3) & ~3;            // multiple of 4 (window is one of 128)
            const bool col_windowed = upper_only || window < n;
            if (slot >= 0) {
                const unsigned* src = saved.bits + (size_t)slot * saved.words;
                if (COMPACT) for (int t = threadIdx.x; t < words; t += THREADS) tab.bits[t] = __ldg(src + t);
                else for (int t = threadIdx.x; t < words; t += THREADS) reinterpret_cast<uint2*>(tab.bits)[t] = make_uint2(__ldg(src + t), 0u);
                __syncthreads();
            } else {
                if (COMPACT) for (int t = threadIdx.x; t < words; t += THREADS) tab.bits[t] = 0u;
                else for (int t = threadIdx.x; t < words; t += THREADS) reinterpret_cast<uint2*>(tab.bits)[t] = make_uint2(0u, 0u);
                __syncthreads();
                // pass 1: occupancy
                expand_row_block<false>(A, B, a_begin, a_end, wl, wh, col_windowed, b_sorted, s_seg, [&](int c, double) {
                    const int o = c - w0;
                    const unsigned m = 1u << (o & 31);
                    unsigned* wp = tab.word_ptr(o >> 5);
                    if (!(*((volatile unsigned*)wp) & m)) atomicOr(wp, m);
                });
                __syncthreads();
            }
            // exclusive popcount prefix (per word, or per group of four words)
            const int units = COMPACT ? words >> 2 : words;
            int nnz_w = 0;
            for (int base = 0; base < units; base += THREADS) {
                const int u = base + threadIdx.x;
                int pc = 0;
                if (u < units) {
                    if (COMPACT) {
                        const uint4 b = reinterpret_cast<const uint4*>(tab.bits)[u];
                        pc = __popc(b.x) + __popc(b.y) + __popc(b.z) + __popc(b.w);
                    } else {
                        pc = __popc(tab.bits[2 * u]);
                    }
                }
                int tot;
                const int ex = block_excl_scan<int>(pc, s_red, &tot);
                if (u < units) {
                    if (COMPACT) tab.pre[u] = (unsigned)(nnz_w + ex); else tab.bits[2 * u + 1] = (unsigned)(nnz_w + ex);
                }
                nnz_w += tot;
            }
            __syncthreads();
            // emit the sorted column indices of this column window; zero the values they will accumulate into
            for (int u = threadIdx.x; u < units; u += THREADS) {
                int64_t pos = out + (COMPACT ? tab.pre[u] : tab.bits[2 * u + 1]);
                const int nw = COMPACT ? 4 : 1;
                for (int k = 0; k < nw; ++k) {
                    const int w = COMPACT ? 4 * u + k : u;
                    unsigned word = *tab.word_ptr(w);
                    while (word) {
                        const int b = __ffs(word) - 1;
                        word &= word - 1;
                        c_idx[pos++] = w0 + (w << 5) + b;
                    }
                }
            }
            double* gv = c_val + out;
            for (int t = threadIdx.x; t < nnz_w; t += THREADS) gv[t] = 0.0;
            __syncthreads();
            // pass 2: values
            expand_row_block<true>(A, B, a_begin, a_end, wl, wh, col_windowed, b_sorted, s_seg,
                                   [&](int c, double v) { atomicAdd(gv + tab.rank(c - w0), v); });
            out += nnz_w;
            __syncthreads();
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
    e = cudaFuncSetAttribute(k_symbolic_bitmap<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_symbolic_bitmap<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 24576);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_numeric_rank<false, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_numeric_rank<true, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 24576);
    return e;
}

// Run up to three independent bin kernels on side streams: fork after everything queued on the main stream so
// far, join back before anything queued after.
struct Fork {
    const LaunchCtx& lc;
    bool used[3] = {false, false, false};
    bool forked = false;
    explicit Fork(const LaunchCtx& l) : lc(l) {}
    cudaStream_t side(int k, bool concurrent) {
        if (!concurrent || lc.aux[k] == nullptr) return lc.stream;
        if (!forked) { cudaEventRecord(lc.fork_ev, lc.stream); forked = true; }
        if (!used[k]) { cudaStreamWaitEvent(lc.aux[k], lc.fork_ev, 0); used[k] = true; }
        return lc.aux[k];
    }
    void join() {
        for (int k = 0; k < 3; ++k)
            if (used[k]) { cudaEventRecord(lc.join_ev[k], lc.aux[k]); cudaStreamWaitEvent(lc.stream, lc.join_ev[k], 0); }
    }
};

static inline int grid_for(int items, int per_block, int cap) {
    int g = (items + per_block - 1) / per_block;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return g;
}

cudaError_t launch_symbolic(const LaunchCtx& lc, const SparseJob& job, const int32_t* d_lists, const int32_t* h_counts,
                            int32_t* d_nnz, int32_t* d_work_counter, const SavedBitmaps& saved) {
    const int up = job.upper_only ? 1 : 0;
    const size_t stride = (size_t)job.nrows;
    const int cap = lc.sm_count * 16;
    const int nonempty = (h_counts[SYM_W64] != 0) + (h_counts[SYM_W256] != 0) + (h_counts[SYM_W1K] != 0) +
                         (h_counts[SYM_BITMAP] != 0);
    Fork fork(lc);
    const bool conc = nonempty > 1;
    if (h_counts[SYM_W64]) {
        constexpr int W = 8;
        k_symbolic_warp<64, W><<<grid_for(h_counts[SYM_W64], W, cap), W * 32, 0, fork.side(0, conc)>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_W64 * stride, h_counts[SYM_W64], d_nnz);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[SYM_W256]) {
        constexpr int W = 8;
        k_symbolic_warp<256, W><<<grid_for(h_counts[SYM_W256], W, cap), W * 32, 0, fork.side(1, conc)>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_W256 * stride, h_counts[SYM_W256], d_nnz);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[SYM_W1K]) {
        constexpr int W = 8;
        k_symbolic_warp<1024, W><<<grid_for(h_counts[SYM_W1K], W, cap), W * 32, 0, fork.side(2, conc)>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_W1K * stride, h_counts[SYM_W1K], d_nnz);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[SYM_BITMAP]) {
        // window = all columns when they fit the per-block shared memory, else the largest multiple of 32 bits
        const size_t max_bits = (g_smem_optin - 24576) * 8;
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
        if (per_sm >= 2)
            k_symbolic_bitmap<512><<<grid_for(h_counts[SYM_BITMAP], 1, lc.sm_count * per_sm), 512, smem, lc.stream>>>(
                job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_BITMAP * stride, h_counts[SYM_BITMAP],
                (int)window_bits, d_nnz, d_work_counter, saved);
        else                                           // one block per SM: make it a full 1024 threads
            k_symbolic_bitmap<1024><<<grid_for(h_counts[SYM_BITMAP], 1, lc.sm_count), 1024, smem, lc.stream>>>(
                job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + SYM_BITMAP * stride, h_counts[SYM_BITMAP],
                (int)window_bits, d_nnz, d_work_counter, saved);
        SB_LAUNCH_CHECK(lc);
    }
    fork.join();
    return cudaSuccess;
}

// words of one saved bitmap: the numeric kernel's rounding of the column count (a multiple of 4 words)
int saved_bitmap_words(int cols) {
    return (int)((((int64_t)cols + 127) & ~(int64_t)127) >> 5);
}

cudaError_t launch_numeric(const LaunchCtx& lc, const SparseJob& job, const int32_t* d_lists, const int32_t* h_counts,
                           const int64_t* c_ptr, int32_t* c_idx, double* c_val, int32_t* d_work_counter,
                           const SavedBitmaps& saved) {
    const int up = job.upper_only ? 1 : 0;
    const size_t stride = (size_t)job.nrows;
    const int cap = lc.sm_count * 16;
    const int nonempty = (h_counts[NUM_W64] != 0) + (h_counts[NUM_W256] != 0) + (h_counts[NUM_W1K] != 0) +
                         (h_counts[NUM_RANK] != 0);
    Fork fork(lc);
    const bool conc = nonempty > 1;
    if (h_counts[NUM_W64]) {
        constexpr int W = 8;
        k_numeric_warp<64, W><<<grid_for(h_counts[NUM_W64], W, cap), W * 32, 0, fork.side(0, conc)>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_W64 * stride, h_counts[NUM_W64], c_ptr,
            c_idx, c_val);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[NUM_W256]) {
        constexpr int W = 8;
        k_numeric_warp<256, W><<<grid_for(h_counts[NUM_W256], W, cap), W * 32, 0, fork.side(1, conc)>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_W256 * stride, h_counts[NUM_W256], c_ptr,
            c_idx, c_val);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[NUM_W1K]) {
        constexpr int W = 4;
        k_numeric_warp<1024, W><<<grid_for(h_counts[NUM_W1K], W, cap), W * 32, 0, fork.side(2, conc)>>>(
            job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_W1K * stride, h_counts[NUM_W1K], c_ptr,
            c_idx, c_val);
        SB_LAUNCH_CHECK(lc);
    }
    if (h_counts[NUM_RANK]) {
        cudaError_t e = cudaMemsetAsync(d_work_counter, 0, sizeof(int32_t), lc.stream);
        if (e != cudaSuccess) return e;
        // ticket -> list position stride: a prime that does not divide the count (so the map is a permutation)
        int ticket_stride = 1;
        for (int cand : {1000003, 999983, 999979, 7919, 104729}) {
            if (h_counts[NUM_RANK] % cand != 0) { ticket_stride = cand % h_counts[NUM_RANK]; break; }
        }
        if (ticket_stride == 0) ticket_stride = 1;
        // pair layout (8 B per 32 columns) while at least two blocks fit an SM; else the compact layout (5 B per
        // 32 columns) with one 1024-thread block per SM and a window of up to 2^20 columns
        const int64_t cols = ((int64_t)job.B.cols + 127) & ~(int64_t)127;
        const size_t two_per_sm = g_smem_optin / 2 - 11264;
        if ((size_t)(cols / 4) <= two_per_sm) {
            const size_t smem = (size_t)(cols / 4);
            int per_sm = (int)((g_smem_optin + 1024) / (smem + 10240));
            if (per_sm > 4) per_sm = 4;
            k_numeric_rank<false, 512><<<grid_for(h_counts[NUM_RANK], 1, lc.sm_count * per_sm), 512, smem, lc.stream>>>(
                job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_RANK * stride, h_counts[NUM_RANK],
                ticket_stride, (int)cols, c_ptr, c_idx, c_val, d_work_counter, saved);
        } else {
            const int64_t window = cols < (1 << 20) ? cols : (1 << 20);
            const size_t smem = (size_t)(window / 8 + window / 32);
            k_numeric_rank<true, 1024><<<grid_for(h_counts[NUM_RANK], 1, lc.sm_count), 1024, smem, lc.stream>>>(
                job.A, job.B, job.row_begin, up, job.d_b_sorted, d_lists + NUM_RANK * stride, h_counts[NUM_RANK],
                ticket_stride, (int)window, c_ptr, c_idx, c_val, d_work_counter, saved);
        }
        SB_LAUNCH_CHECK(lc);
    }
    fork.join();
    return cudaSuccess;
}

}  // namespace sb
