// spgemm_dense.cu -- sparse x sparse -> dense row-major float64.
//
// Replaces dense_nosym / dense_sym (/root/reference/src/sparse_sparse_dense.cpp:79-131, :13-74): the reference
// callocs C and scatter-adds every product into it.  Here the zero fill and the scatter are one pass: a thread
// block owns (row, column tile), builds the tile in shared memory and streams it out with 128-bit stores, so
// every byte of C is written exactly once and never read.  Tiles that receive no product skip shared memory.
#include <cstdlib>

#include "internal.h"

namespace sb {

#define SB_LAUNCH_CHECK(lc)                  \
    do {                                     \
        ++*(lc).launches;                    \
        cudaError_t e_ = cudaGetLastError(); \
        if (e_ != cudaSuccess) return e_;    \
    } while (0)

constexpr int kDenseThreads = 512;
constexpr int kDenseTileMax = 12288;   // doubles per tile: 96 KB (+ 8 KB scratch), two blocks per SM

// Write `count` doubles from shared `src` (or zeros when src == nullptr) to global `dst` with 16-byte stores
// where alignment allows.
__device__ __forceinline__ void stream_out(double* __restrict__ dst, const double* src, int count) {
    if (count <= 0) return;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int head = (int)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1);   // 1: dst is 8- but not 16-byte aligned
    if (head && tid == 0 && count > 0) st_stream_f64(dst, src ? src[0] : 0.0);
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

template <bool UPPER>
__global__ void __launch_bounds__(kDenseThreads)
k_dense_tiles(Csr A, Csr B, const int32_t* __restrict__ b_sorted_flag, int row_begin, int nrows, int tile_w,
              int ntiles, double* __restrict__ C) {
    extern __shared__ double acc[];
    __shared__ SegScratch<kDenseThreads> s_seg;
    const int n = B.cols;
    const bool b_sorted = *b_sorted_flag != 0;
    for (int64_t item = blockIdx.x; item < (int64_t)nrows * ntiles; item += gridDim.x) {
        const int r = (int)(item / ntiles), t = (int)(item % ntiles);
        const int i = row_begin + r;
        const int t0 = t * tile_w, t1 = min(n, t0 + tile_w);
        const int lo = UPPER ? max(t0, i) : t0;
        const int a_begin = __ldg(A.ptr + i), a_end = __ldg(A.ptr + i + 1);
        double* out = C + (size_t)r * n + t0;
        if (a_begin == a_end || lo >= t1) {
            stream_out(out, nullptr, t1 - t0);
            continue;
        }
        for (int x = threadIdx.x; x < t1 - t0; x += blockDim.x) acc[x] = 0.0;
        __syncthreads();
        expand_row_block<true>(A, B, a_begin, a_end, lo, t1, UPPER || ntiles > 1, b_sorted, s_seg,
                               [&](int c, double v) { atomicAdd(acc + (c - t0), v); });
        __syncthreads();
        stream_out(out, acc, t1 - t0);
        __syncthreads();
    }
}

// Variant for outputs that are mostly zeros (few products per output element): a block owns a whole row of C.
// It issues the gathers for the row of A, streams zeros over the row while they are in flight, and then adds
// the products into the row with fire-and-forget float64 reductions that resolve in L2, where the freshly
// written line still sits -- so DRAM sees each byte of C once, and no shared-memory tile is needed.
constexpr int kDenseRedThreads = 256;

// The row's products are gathered FIRST, into a small shared-memory list; only then are the zeros streamed and
// the staged products added.  Gathering first matters: loads issued after the 160 KB of zero stores queue behind
// them in the SM's memory pipeline, and by the time the reductions were issued the freshly written lines had
// left L2 (every reduction then cost a DRAM read-modify-write).  Rows with more products than the list holds
// zero first and reduce directly.
constexpr int kDenseStage = 1024;

template <bool UPPER>
__global__ void __launch_bounds__(kDenseRedThreads)
k_dense_rows_red(Csr A, Csr B, const int32_t* __restrict__ b_sorted_flag, int row_begin, int nrows,
                 double* __restrict__ C) {
    __shared__ SegScratch<kDenseRedThreads> s_seg;
    __shared__ double s_val[kDenseStage];
    __shared__ int s_col[kDenseStage];
    __shared__ int s_count;
    const int n = B.cols;
    const bool b_sorted = *b_sorted_flag != 0;
    for (int r = blockIdx.x; r < nrows; r += gridDim.x) {
        const int i = row_begin + r;
        double* row = C + (size_t)r * n;
        const int a_begin = __ldg(A.ptr + i), a_end = __ldg(A.ptr + i + 1);
        if (threadIdx.x == 0) s_count = 0;
        __syncthreads();
        if (a_begin != a_end)
            expand_row_block<true>(A, B, a_begin, a_end, UPPER ? i : 0, n, UPPER, b_sorted, s_seg,
                                   [&](int c, double v) {
                                       const int pos = atomicAdd(&s_count, 1);
                                       if (pos < kDenseStage) { s_col[pos] = c; s_val[pos] = v; }
                                   });
        stream_out(row, nullptr, n);            // evict-first: keeps A and B resident in L2 under the write stream
        __syncthreads();                        // zero stores ordered before the reductions; list complete
        const int staged = s_count;
        if (staged <= kDenseStage) {
            for (int t = threadIdx.x; t < staged; t += blockDim.x)
                atomicAdd(row + s_col[t], s_val[t]);
        } else {
            // too many products to stage: enumerate them again, reducing directly
            expand_row_block<true>(A, B, a_begin, a_end, UPPER ? i : 0, n, UPPER, b_sorted, s_seg,
                                   [&](int c, double v) { atomicAdd(row + c, v); });
        }
        __syncthreads();
    }
}

// C[j,i] = C[i,j] for j > i (n x n, in place): 32x32 tiles through shared memory so both the read of the
// upper tile and the write of its mirror image are coalesced.
__global__ void __launch_bounds__(256)
k_mirror_upper(double* __restrict__ C, int n) {
    __shared__ double tile[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    for (int y = ty; y < 32; y += 8) {
        const int i = bi * 32 + y, j = bj * 32 + tx;
        tile[y][tx] = (i < n && j < n) ? C[(size_t)i * n + j] : 0.0;
    }
    __syncthreads();
    for (int y = ty; y < 32; y += 8) {
        // destination element (row = bj*32 + y, col = bi*32 + tx) takes source (bi*32 + tx, bj*32 + y)
        const int row = bj * 32 + y, col = bi * 32 + tx;
        if (row < n && col < n && row > col) C[(size_t)row * n + col] = tile[tx][y];
    }
}

// C = C + C^T - diag(C) in place (what the reference's compute_full_matrix=1 produces from the full product).
__global__ void __launch_bounds__(256)
k_symmetrize_add(double* __restrict__ C, int n) {
    __shared__ double up[32][33], dn[32][33];
    const int bi = blockIdx.y, bj = blockIdx.x;
    if (bj < bi) return;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int y = ty; y < 32; y += 8) {
        const int i = bi * 32 + y, j = bj * 32 + tx;
        up[y][tx] = (i < n && j < n) ? C[(size_t)i * n + j] : 0.0;
        const int i2 = bj * 32 + y, j2 = bi * 32 + tx;
        dn[y][tx] = (i2 < n && j2 < n) ? C[(size_t)i2 * n + j2] : 0.0;
    }
    __syncthreads();
    for (int y = ty; y < 32; y += 8) {
        const int i = bi * 32 + y, j = bj * 32 + tx;
        if (i < n && j < n && (bi != bj || i != j)) {
            if (bi != bj || j > i) C[(size_t)i * n + j] = up[y][tx] + dn[tx][y];
        }
        const int i2 = bj * 32 + y, j2 = bi * 32 + tx;
        if (i2 < n && j2 < n && (bi != bj ? true : i2 > j2)) C[(size_t)i2 * n + j2] = dn[y][tx] + up[tx][y];
    }
}

static size_t g_dense_smem_optin = 0;

cudaError_t dense_kernels_configure() {
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    g_dense_smem_optin = (size_t)optin;
    e = cudaFuncSetAttribute(k_dense_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_dense_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - 20480);
}

cudaError_t launch_dense(const LaunchCtx& lc, const Csr& A, const Csr& B, const int32_t* d_b_sorted, bool upper_only,
                         int row_begin, int nrows, double* d_c, int mode, double products_per_out) {
    const int n = B.cols;
    if (nrows <= 0 || n <= 0) return cudaSuccess;
    if (mode == 0) mode = products_per_out < 0.5 ? 2 : 1;
    if (mode == 2) {
        const int grid = nrows;              // one block per row: the hardware scheduler keeps every SM writing
        if (upper_only)
            k_dense_rows_red<true><<<grid, kDenseRedThreads, 0, lc.stream>>>(A, B, d_b_sorted, row_begin, nrows, d_c);
        else
            k_dense_rows_red<false><<<grid, kDenseRedThreads, 0, lc.stream>>>(A, B, d_b_sorted, row_begin, nrows, d_c);
        SB_LAUNCH_CHECK(lc);
        return cudaSuccess;
    }
    // equal-width column tiles, each <= kDenseTileMax doubles and a multiple of 2 doubles wide
    int ntiles = (n + kDenseTileMax - 1) / kDenseTileMax;
    int tile_w = (n + ntiles - 1) / ntiles;
    tile_w = (tile_w + 1) & ~1;
    ntiles = (n + tile_w - 1) / tile_w;
    const size_t smem = (size_t)tile_w * sizeof(double);
    const int64_t items = (int64_t)nrows * ntiles;
    int per_sm = (int)(g_dense_smem_optin / (smem + 10240));
    if (per_sm > 4) per_sm = 4;
    if (per_sm < 1) per_sm = 1;
    // persistent-style grid: a multiple of the SM count, several waves so late rows balance
    int64_t grid = (int64_t)lc.sm_count * per_sm * 8;
    if (grid > items) grid = items;
    if (upper_only)
        k_dense_tiles<true><<<(unsigned)grid, kDenseThreads, smem, lc.stream>>>(A, B, d_b_sorted, row_begin, nrows,
                                                                                 tile_w, ntiles, d_c);
    else
        k_dense_tiles<false><<<(unsigned)grid, kDenseThreads, smem, lc.stream>>>(A, B, d_b_sorted, row_begin, nrows,
                                                                                  tile_w, ntiles, d_c);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

cudaError_t launch_mirror(const LaunchCtx& lc, double* d_c, int n) {
    if (n <= 1) return cudaSuccess;
    const int nb = (n + 31) / 32;
    k_mirror_upper<<<dim3(nb, nb), 256, 0, lc.stream>>>(d_c, n);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

cudaError_t launch_symmetrize(const LaunchCtx& lc, double* d_c, int n) {
    if (n <= 1) return cudaSuccess;
    const int nb = (n + 31) / 32;
    k_symmetrize_add<<<dim3(nb, nb), 256, 0, lc.stream>>>(d_c, n);
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

}  // namespace sb
