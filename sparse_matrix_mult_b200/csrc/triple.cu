// triple.cu -- fused triple product C = H Q H^T into a dense row-major float64 matrix.
//
// Replaces triple_product (/root/reference/src/sparse_sparse_dense.cpp:141-249).  The reference expands
// t = H[i,:] Q into a dense scratch row and then takes a sparse dot of t with EVERY row r >= i of H
// (n(n+1)/2 * nnz/row gathers).  Here a thread block owns (row i, column tile): it streams the products
// w = h_ij * q_jc of the expansion (never stored) and contracts each against row c of H^T, scatter-adding
// w * h_rc into the tile of C[i, :] held in shared memory; the tile then leaves with 128-bit streaming
// stores.  Work is P1 + P2 instead of n^2 * nnz/row, H Q is never materialised, and C is
// written once.
#include "internal.h"

namespace sb {

#define SB_LAUNCH_CHECK(lc)                  \
    do {                                     \
        ++*(lc).launches;                    \
        cudaError_t e_ = cudaGetLastError(); \
        if (e_ != cudaSuccess) return e_;    \
    } while (0)

constexpr int kTripleThreads = 512;
constexpr int kTripleTileMax = 12288;

__device__ __forceinline__ void triple_stream_out(double* __restrict__ dst, const double* src, int count) {
    if (count <= 0) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int head = (int)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1);
    if (head && tid == 0) st_stream_f64(dst, src ? src[0] : 0.0);
    const int pairs = (count - head) >> 1;
    double* d2 = dst + head;
    if (src) {
        const double* s2 = src + head;
        for (int t = tid; t < pairs; t += nt) st_stream_f64x2(d2 + 2 * t, s2[2 * t], s2[2 * t + 1]);
    } else {
        for (int t = tid; t < pairs; t += nt) st_stream_f64x2(d2 + 2 * t, 0.0, 0.0);
    }
    const int tail = head + 2 * pairs;
    if (tail < count && tid == nt - 1) st_stream_f64(dst + tail, src ? src[tail] : 0.0);
}

// counters[0] += expansion products (P1), counters[1] += scatter-adds performed (P2); counters[2] = row ticket
__device__ __forceinline__ void triple_flush_counters(unsigned long long p1, unsigned long long p2,
                                                      unsigned long long* s_cnt, unsigned long long* counters) {
    if (!counters) return;
    p1 = warp_sum(p1);
    p2 = warp_sum(p2);
    if (lane_id() == 0) {
        if (p1) atomicAdd(&s_cnt[0], p1);
        if (p2) atomicAdd(&s_cnt[1], p2);
    }
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(counters + threadIdx.x, s_cnt[threadIdx.x]);
}

template <bool UPPER>
__global__ void __launch_bounds__(kTripleThreads)
k_triple_tiles(Csr H, Csr Q, Csr Ht, const int32_t* __restrict__ ht_desc, int row_begin, int nrows, int tile_w,
               int ntiles, double* __restrict__ C, unsigned long long* __restrict__ counters) {
    extern __shared__ double acc[];
    __shared__ unsigned long long s_cnt[2];
    __shared__ SegScratch<kTripleThreads> s_seg;
    const int n = H.rows;
    const bool desc = ht_desc != nullptr && *ht_desc != 0;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long p1 = 0, p2 = 0;
    for (int64_t item = blockIdx.x; item < (int64_t)nrows * ntiles; item += gridDim.x) {
        const int r = (int)(item / ntiles), t = (int)(item % ntiles);
        const int i = row_begin + r;
        const int t0 = t * tile_w, t1 = min(n, t0 + tile_w);
        const int lo = UPPER ? max(t0, i) : t0;
        const int first_t = UPPER ? i / tile_w : 0;
        const int h_begin = __ldg(H.ptr + i), h_end = __ldg(H.ptr + i + 1);
        double* out = C + (size_t)r * n + t0;
        if (h_begin == h_end || lo >= t1) {
            triple_stream_out(out, nullptr, t1 - t0);
            continue;
        }
        for (int x = threadIdx.x; x < t1 - t0; x += blockDim.x) acc[x] = 0.0;
        __syncthreads();
        expand_row_block<true>(H, Q, h_begin, h_end, 0, 0, false, false, s_seg, [&](int c, double w) {
            if (t == first_t) ++p1;                 // count the expansion once per row, not per tile
            const int s = __ldg(Ht.ptr + c), e = __ldg(Ht.ptr + c + 1);
            for (int q = s; q < e; ++q) {
                const int k = __ldg(Ht.idx + q);
                if (k < lo) { if (desc) break; else continue; }     // descending rows: nothing useful follows
                if (k < t1) {
                    atomicAdd(acc + (k - t0), w * __ldg(Ht.val + q));
                    ++p2;
                }
            }
        });
        __syncthreads();
        triple_stream_out(out, acc, t1 - t0);
        __syncthreads();
    }
    triple_flush_counters(p1, p2, s_cnt, counters);
}

// Variant without a shared-memory tile: the block owns a whole row of C, streams zeros over it while the
// gathers of H[i,:] are in flight, then adds every contribution with a float64 reduction that resolves in L2
// (native RED.ADD.F64 -- the shared-memory path needs a compare-and-swap loop per add).
// The rows being accumulated must stay L2 resident (a reduction that misses L2 is a DRAM read-modify-write, ~8x
// slower): the grid is persistent and sized so that rows-in-flight x 8n bytes fits a share of the 126 MB L2;
// wide outputs therefore run one 1024-thread block per SM instead of eight 256-thread blocks.
template <bool UPPER, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_triple_rows_red(Csr H, Csr Q, Csr Ht, const int32_t* __restrict__ ht_desc, int row_begin, int nrows,
                  double* __restrict__ C, unsigned long long* __restrict__ counters) {
    __shared__ unsigned long long s_cnt[2];
    __shared__ SegScratch<THREADS> s_seg;
    const int n = H.rows;
    const bool desc = ht_desc != nullptr && *ht_desc != 0;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long p1 = 0, p2 = 0;
    __shared__ int s_row;
    while (true) {
        // rows are handed out in order through counters[2] (upper-triangle rows get cheaper towards the bottom,
        // so in-order dynamic scheduling is longest-first)
        if (threadIdx.x == 0) s_row = (int)atomicAdd(counters + 2, 1ULL);
        __syncthreads();
        const int r = s_row;
        __syncthreads();
        if (r >= nrows) break;
        const int i = row_begin + r;
        const int lo = UPPER ? i : 0;
        double* row = C + (size_t)r * n;
        triple_stream_out(row, nullptr, n);
        expand_row_block<true>(H, Q, __ldg(H.ptr + i), __ldg(H.ptr + i + 1), 0, 0, false, false, s_seg,
                               [&](int c, double w) {
                                   ++p1;
                                   const int s = __ldg(Ht.ptr + c), e = __ldg(Ht.ptr + c + 1);
                                   for (int q = s; q < e; ++q) {
                                       const int k = __ldg(Ht.idx + q);
                                       if (k < lo) { if (desc) break; else continue; }
                                       atomicAdd(row + k, w * __ldg(Ht.val + q));
                                       ++p2;
                                   }
                               });
    }
    triple_flush_counters(p1, p2, s_cnt, counters);
}

static size_t g_triple_smem_optin = 0;

cudaError_t triple_kernels_configure() {
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    g_triple_smem_optin = (size_t)optin;
    e = cudaFuncSetAttribute(k_triple_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_triple_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
}

cudaError_t launch_triple(const LaunchCtx& lc, const Csr& H, const Csr& Q, const Csr& Ht, const int32_t* d_ht_desc,
                          bool upper_only,
                          int row_begin, int nrows, double* d_c, unsigned long long* d_counters, int mode) {
    const int n = H.rows;
    if (nrows <= 0 || n <= 0) return cudaSuccess;
    if (mode == 0) mode = 2;
    if (mode == 2) {
        const double l2_budget = 64.0e6;                                  // bytes of C rows in flight
        const double row_bytes = 8.0 * n * (upper_only ? 0.6 : 1.0);      // upper rows touch [i, n) only
        int rows_in_flight = (int)(l2_budget / row_bytes);
        if (rows_in_flight < lc.sm_count) rows_in_flight = lc.sm_count;
        const bool big = rows_in_flight < lc.sm_count * 4;                // few rows allowed: fat blocks
        const int per_sm = big ? 1 : (rows_in_flight / lc.sm_count > 8 ? 8 : rows_in_flight / lc.sm_count);
        int grid = lc.sm_count * per_sm;
        if (grid > nrows) grid = nrows;
        if (big) {
            if (upper_only)
                k_triple_rows_red<true, 1024><<<grid, 1024, 0, lc.stream>>>(H, Q, Ht, d_ht_desc, row_begin, nrows, d_c, d_counters);
            else
                k_triple_rows_red<false, 1024><<<grid, 1024, 0, lc.stream>>>(H, Q, Ht, d_ht_desc, row_begin, nrows, d_c, d_counters);
        } else {
            if (upper_only)
                k_triple_rows_red<true, 256><<<grid, 256, 0, lc.stream>>>(H, Q, Ht, d_ht_desc, row_begin, nrows, d_c, d_counters);
            else
                k_triple_rows_red<false, 256><<<grid, 256, 0, lc.stream>>>(H, Q, Ht, d_ht_desc, row_begin, nrows, d_c, d_counters);
        }
        SB_LAUNCH_CHECK(lc);
        return cudaSuccess;
    }
    int ntiles = (n + kTripleTileMax - 1) / kTripleTileMax;
    int tile_w = (n + ntiles - 1) / ntiles;
    tile_w = (tile_w + 1) & ~1;
    ntiles = (n + tile_w - 1) / tile_w;
    const size_t smem = (size_t)tile_w * sizeof(double);
    const int64_t items = (int64_t)nrows * ntiles;
    int per_sm = (int)(g_triple_smem_optin / (smem + 10240));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)lc.sm_count * per_sm * 8;
    if (grid > items) grid = items;
    if (upper_only)
        k_triple_tiles<true><<<(unsigned)grid, kTripleThreads, smem, lc.stream>>>(H, Q, Ht, d_ht_desc, row_begin, nrows, tile_w,
                                                                                   ntiles, d_c, d_counters);
    else
        k_triple_tiles<false><<<(unsigned)grid, kTripleThreads, smem, lc.stream>>>(H, Q, Ht, d_ht_desc, row_begin, nrows, tile_w,
                                                                                    ntiles, d_c, d_counters);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

}  // namespace sb
