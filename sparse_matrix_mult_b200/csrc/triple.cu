// triple.cu -- fused triple product C = H Q H^T into a dense row-major float64 matrix.
//
// Replaces triple_product (/root/reference/src/sparse_sparse_dense.cpp:141-249).  The reference expands
// t = H[i,:] Q into a dense scratch row (:187-198) and then takes a sparse dot of t with EVERY row k >= i of H
// (:201-216, n(n+1)/2 * nnz/row gathers).  Here a thread block owns row i of C: it streams the products
// w = h_ij * q_jc of the expansion (never stored) and contracts each against row c of H^T,
// C[i,k] += w * h_kc for k >= i.  Work is P1 + P2 instead of n^2/2 * nnz/row, H Q is never
// materialised, and every byte of C is written once.
//
// k_triple_window (default): the row's accumulator lives in SHARED memory -- a window of up to ~26,800 doubles
// starting at the diagonal; columns beyond the window (only the first rows of very wide outputs have any) are
// accumulated with float64 reductions that resolve in L2.  Shared-memory accumulation runs at 2-3x the chip-wide
// L2 reduction rate (scripts/micro/atomic_bw.cu: 540 vs 197 G adds/s), needs no L2 residency of the C rows in
// flight, and the finished window leaves with coalesced 128-bit streaming stores.
// k_triple_rows_red (round 1, kept selectable for A/B runs: SPGEMM_B200_TRIPLE_MODE=2): every add is an L2 reduction.
#include <cstdlib>

#include "internal.h"

namespace sb {

#define SB_LAUNCH_CHECK(lc)                  \
    do {                                     \
        ++*(lc).launches;                    \
        cudaError_t e_ = cudaGetLastError(); \
        if (e_ != cudaSuccess) return e_;    \
    } while (0)

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

// ---------------------------------------------------------------------------------------------------
// Shared-memory window kernel.
//
// Per row i of C:
//   1. up to blockDim entries (j, h_ij) of H[i,:] are loaded one per thread together with the extent of row j of Q;
//      a block-wide prefix sum numbers the products of the expansion 0..total-1;
//   2. the products are taken 32 at a time by the warps, round-robin IN ORDER (chunk k goes to warp k mod nwarp):
//      lane l of a warp takes product 32k + l, finds its row of Q in the prefix table, loads (c, q_jc) -- coalesced,
//      a row of Q is contiguous, and streamed past L2 (evict-first: a row of Q is used once per row of C) -- and
//      the extent of row c of H^T;
//   3. contraction as ONE flat stream: the rows of H^T of the warp's 32 products are numbered by a warp prefix sum
//      of their lengths and the warp walks the concatenation 32 entries per step.  The owner of an entry is found
//      from a 32-bit mask of the row starts inside the step (one warp-wide OR reduction + popcount); its weight
//      w = h_ij * q_jc and the offset of its row come from a 32-entry per-warp table in shared memory.  When
//      consecutive products have consecutive c (banded Q) their rows of H^T are adjacent in memory, so every load
//      of the stream is fully coalesced; for any other Q it is a balanced gather.  Every lane is active whatever
//      the row lengths of H^T (the round-1 kernel walked one row per thread: 17.8 of 32 lanes active);
//   4. entries with k >= i (upper mode) are added to the shared window [w0, w1) (float64 CAS loop) or, beyond it,
//      to C itself (L2 reduction);
//   5. the window is streamed out and cleared in one pass.
//
// SYNC (cooperative launch): the grid processes the rows in batches of gridDim.x, one row per block, with a grid
// barrier between batches.  Rows of H are sorted by column and step 2 walks them in order, so all blocks sweep the
// column range of H^T together: at any moment the chip reads a narrow band of H^T (cfg 5: ~20 MB of the 96 MB),
// which stays L2 resident, instead of gathering from all of it (round 1: 36 GB of DRAM reads, L2 hit rate 50 %).
// Without SYNC rows are handed out through an atomic ticket.
//
// Dynamic shared memory: acc[win_cap] | hv[nt] | w[nt] | qs[nt] | pre[nt + 1] | d[nt].
struct TripleScratch {
    unsigned red[33];
    unsigned long long cnt[2];
    int row;
};
__host__ __device__ inline size_t triple_window_smem(int win_cap, int threads) {
    return (size_t)win_cap * 8 + (size_t)threads * 28 + 16;
}

__device__ __forceinline__ int ld_stream_i32(const int32_t* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream_f64(const double* p) { return __ldcs(p); }

// counters[3]: monotone arrival counter of the grid barrier (zeroed before the launch)
__device__ __forceinline__ void grid_barrier(unsigned long long* counter, unsigned long long target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1ULL);
        while (*((volatile unsigned long long*)counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <bool UPPER, bool SYNC>
__global__ void __launch_bounds__(1024, 1)
k_triple_window(Csr H, Csr Q, Csr Ht, int row_begin, int nrows, int win_cap,
                double* __restrict__ C, unsigned long long* __restrict__ counters) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ TripleScratch S;
    const int n = H.rows;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    double* acc = reinterpret_cast<double*>(s_raw);
    double* s_hv = acc + win_cap;                                   // win_cap is even: 16-byte aligned
    double* s_w = s_hv + nt + warp * 32;                            // this warp's 32 weights
    int* s_qs = reinterpret_cast<int*>(s_hv + 2 * nt);
    unsigned* s_pre = reinterpret_cast<unsigned*>(s_qs + nt);
    int* s_d = reinterpret_cast<int*>(s_pre + nt + 1) + warp * 32;  // this warp's 32 row offsets
    const unsigned lt_mask = (1u << lane) - 1u, le_mask = lt_mask | (1u << lane);
    if (tid < 2) S.cnt[tid] = 0;
    for (int t = tid; t < win_cap; t += nt) acc[t] = 0.0;
    __syncthreads();
    unsigned long long p1 = 0;
    unsigned p2 = 0;
    unsigned long long p2_total = 0;
    const int nbatch = SYNC ? (nrows + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    for (int batch = 0; ; ++batch) {
        int r;
        if (SYNC) {
            if (batch >= nbatch) break;
            r = batch * (int)gridDim.x + (int)blockIdx.x;
        } else {
            if (tid == 0) S.row = (int)atomicAdd(counters + 2, 1ULL);
            __syncthreads();
            r = S.row;
            __syncthreads();
            if (r >= nrows) break;
        }
        if (r < nrows) {
            const int i = row_begin + r;
            const int lo = UPPER ? i : 0;
            const int w0 = lo, w1 = min(n, lo + win_cap);
            double* row = C + (size_t)r * n;
            const int h_begin = __ldg(H.ptr + i), h_end = __ldg(H.ptr + i + 1);
            // zeros left of the diagonal (never touched again: evict-first) and right of the window (about to take
            // reductions: default policy so the lines stay in L2)
            triple_stream_out(row, nullptr, w0);
            for (int t = w1 + tid; t < n; t += nt) row[t] = 0.0;
            for (int base = h_begin; base < h_end; base += nt) {
                const int cnt = min(nt, h_end - base);
                unsigned len = 0;
                if (tid < cnt) {
                    const int j = __ldg(H.idx + base + tid);
                    const int qs = __ldg(Q.ptr + j);
                    len = (unsigned)(__ldg(Q.ptr + j + 1) - qs);
                    s_hv[tid] = __ldg(H.val + base + tid);
                    s_qs[tid] = qs;
                }
                unsigned total;
                const unsigned ex = block_excl_scan<unsigned>(len, S.red, &total);
                if (tid < cnt) s_pre[tid] = ex;
                if (tid == 0) s_pre[cnt] = total;
                __syncthreads();                           // tables complete; tail zeros ordered before reductions
                if (tid == 0) p1 += total;
                for (unsigned f0 = (unsigned)warp * 32u; f0 < total; f0 += (unsigned)nwarp * 32u) {
                    const unsigned f = f0 + lane;
                    const bool valid = f < total;
                    int hs = 0;
                    unsigned hlen = 0;
                    double w = 0.0;
                    if (valid) {
                        int a = 0, b = cnt;                // largest a with pre[a] <= f  (pre[cnt] = total > f)
                        while (b - a > 1) {
                            const int mid = (a + b) >> 1;
                            if (s_pre[mid] <= f) a = mid; else b = mid;
                        }
                        const int q = s_qs[a] + (int)(f - s_pre[a]);
                        const int c = ld_stream_i32(Q.idx + q);
                        w = s_hv[a] * ld_stream_f64(Q.val + q);
                        hs = __ldg(Ht.ptr + c);
                        hlen = (unsigned)(__ldg(Ht.ptr + c + 1) - hs);
                    }
                    // flat numbering of the entries of the 32 rows of H^T
                    const unsigned incl = warp_incl_scan(hlen);
                    const unsigned pre = incl - hlen;
                    const unsigned L = __shfl_sync(FULL, incl, 31);
                    const unsigned nonempty = __ballot_sync(FULL, hlen > 0);
                    if (hlen > 0) {
                        const int rank = __popc(nonempty & lt_mask);
                        s_w[rank] = w;
                        s_d[rank] = hs - (int)pre;
                    }
                    __syncwarp();
                    int heads_before = 0;
                    for (unsigned eb = 0; eb < L; eb += 32) {
                        const unsigned bit = (hlen > 0 && pre >= eb && pre < eb + 32) ? 1u << (pre - eb) : 0u;
                        const unsigned heads = __reduce_or_sync(FULL, bit);
                        const int owner = heads_before + __popc(heads & le_mask) - 1;
                        heads_before += __popc(heads);
                        const unsigned p = eb + lane;
                        if (p < L) {
                            const int addr = (int)p + s_d[owner];
                            const int k = __ldg(Ht.idx + addr);
                            if (k >= lo) {
                                const double x = s_w[owner] * __ldg(Ht.val + addr);
                                if (k < w1) atomicAdd(acc + (k - w0), x); else atomicAdd(row + k, x);
                                ++p2;
                            }
                        }
                    }
                    __syncwarp();                          // the per-warp tables are rewritten by the next chunk
                }
                __syncthreads();                           // tables are rewritten by the next slice / row
            }
            p2_total += p2;
            p2 = 0;
            // window out (coalesced 128-bit streaming stores) and cleared for the next row
            const int count = w1 - w0;
            double* dst = row + w0;
            const int head = (int)((reinterpret_cast<uintptr_t>(dst) >> 3) & 1);
            if (head && tid == 0 && count > 0) { st_stream_f64(dst, acc[0]); acc[0] = 0.0; }
            const int pairs = (count - head) >> 1;
            double* d2 = dst + head;
            double* s2 = acc + head;
            for (int t = tid; t < pairs; t += nt) {
                st_stream_f64x2(d2 + 2 * t, s2[2 * t], s2[2 * t + 1]);
                s2[2 * t] = 0.0;
                s2[2 * t + 1] = 0.0;
            }
            const int tail = head + 2 * pairs;
            if (tail < count && tid == nt - 1) { st_stream_f64(dst + tail, acc[tail]); acc[tail] = 0.0; }
        }
        if (SYNC) grid_barrier(counters + 3, (unsigned long long)(batch + 1) * gridDim.x);
        else __syncthreads();                              // clearing ordered before the next row's adds
    }
    triple_flush_counters(p1, p2_total, S.cnt, counters);
}

// ---------------------------------------------------------------------------------------------------
// Round-1 kernel: the block owns a whole row of C in global memory, streams zeros over it and adds every
// contribution with a float64 reduction that resolves in L2.  The rows being accumulated must stay L2 resident
// (a reduction that misses L2 is a DRAM read-modify-write), so the grid is persistent and sized so that
// rows-in-flight x 8n bytes fits a share of the 126 MB L2.
template <bool UPPER, int THREADS>
__global__ void __launch_bounds__(THREADS)
k_triple_rows_red(Csr H, Csr Q, Csr Ht, int ht_desc, int row_begin, int nrows,
                  double* __restrict__ C, unsigned long long* __restrict__ counters) {
    __shared__ unsigned long long s_cnt[2];
    __shared__ SegScratch<THREADS> s_seg;
    const int n = H.rows;
    const bool desc = ht_desc != 0;
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long p1 = 0, p2 = 0;
    __shared__ int s_row;
    while (true) {
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

// ---------------------------------------------------------------------------------------------------
static size_t g_triple_smem_optin = 0, g_triple_smem_sm = 0;

template <bool UPPER, bool SYNC>
static cudaError_t configure_window() {
    return cudaFuncSetAttribute(k_triple_window<UPPER, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)(g_triple_smem_optin - sizeof(TripleScratch) - 64));
}

cudaError_t triple_kernels_configure() {
    int dev = 0, optin = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&per_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
    if (e != cudaSuccess) return e;
    g_triple_smem_optin = (size_t)optin;
    g_triple_smem_sm = (size_t)per_sm;
    if ((e = configure_window<true, true>()) != cudaSuccess) return e;
    if ((e = configure_window<true, false>()) != cudaSuccess) return e;
    if ((e = configure_window<false, true>()) != cudaSuccess) return e;
    return configure_window<false, false>();
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

template <bool UPPER, bool SYNC>
static cudaError_t launch_window(int grid, int threads, size_t smem, cudaStream_t st, Csr H, Csr Q, Csr Ht, int row_begin,
                                 int nrows, int win_cap, double* d_c, unsigned long long* d_counters) {
    if (!SYNC) {
        k_triple_window<UPPER, false><<<grid, threads, smem, st>>>(H, Q, Ht, row_begin, nrows, win_cap, d_c, d_counters);
        return cudaGetLastError();
    }
    void* args[] = {&H, &Q, &Ht, &row_begin, &nrows, &win_cap, &d_c, &d_counters};
    return cudaLaunchCooperativeKernel((const void*)k_triple_window<UPPER, true>, dim3(grid), dim3(threads), args, smem, st);
}

cudaError_t launch_triple(const LaunchCtx& lc, const Csr& H, const Csr& Q, const Csr& Ht, bool ht_desc,
                          bool upper_only, int row_begin, int nrows, double* d_c, unsigned long long* d_counters,
                          int64_t ht_nnz, int mode) {
    const int n = H.rows;
    if (nrows <= 0 || n <= 0) return cudaSuccess;
    if (mode == 2) {
        const double l2_budget = 64.0e6;                                  // bytes of C rows in flight
        const double row_bytes = 8.0 * n * (upper_only ? 0.6 : 1.0);      // upper rows touch [i, n) only
        int rows_in_flight = (int)(l2_budget / row_bytes);
        if (rows_in_flight < lc.sm_count) rows_in_flight = lc.sm_count;
        const bool big = rows_in_flight < lc.sm_count * 4;                // few rows allowed: fat blocks
        const int per_sm = big ? 1 : (rows_in_flight / lc.sm_count > 8 ? 8 : rows_in_flight / lc.sm_count);
        int grid = lc.sm_count * per_sm;
        if (grid > nrows) grid = nrows;
        const int d = ht_desc ? 1 : 0;
        if (big) {
            if (upper_only)
                k_triple_rows_red<true, 1024><<<grid, 1024, 0, lc.stream>>>(H, Q, Ht, d, row_begin, nrows, d_c, d_counters);
            else
                k_triple_rows_red<false, 1024><<<grid, 1024, 0, lc.stream>>>(H, Q, Ht, d, row_begin, nrows, d_c, d_counters);
        } else {
            if (upper_only)
                k_triple_rows_red<true, 256><<<grid, 256, 0, lc.stream>>>(H, Q, Ht, d, row_begin, nrows, d_c, d_counters);
            else
                k_triple_rows_red<false, 256><<<grid, 256, 0, lc.stream>>>(H, Q, Ht, d, row_begin, nrows, d_c, d_counters);
        }
        SB_LAUNCH_CHECK(lc);
        return cudaSuccess;
    }
    // ---- shared-memory window kernel ----
    // window: the widest row segment this launch accumulates (row `row_begin` in upper mode), capped by what one
    // block may hold; 1, 2 or 4 blocks per SM of 1024 / 512 / 256 threads (32 warps per SM at <= 64 registers)
    const size_t fixed = sizeof(TripleScratch) + 1024;                    // static scratch + per-block reserve
    const int cap_max = (int)(((g_triple_smem_optin - sizeof(TripleScratch) - 64 - triple_window_smem(0, 1024)) / 8) & ~(size_t)1);
    int need = upper_only ? n - row_begin : n;
    need = (need + 1) & ~1;
    if (need < 2) need = 2;
    int win = need < cap_max ? need : cap_max;
    if (const int o = env_int("SPGEMM_B200_TRIPLE_WIN", 0))              // experiments: force the window / residency
        win = (o < cap_max ? o : cap_max) & ~1;
    int per_sm = 1;
    for (int cand : {4, 2}) {
        if (g_triple_smem_sm / (triple_window_smem(win, 1024 / cand) + fixed) >= (size_t)cand) { per_sm = cand; break; }
    }
    const int threads = 1024 / per_sm;
    int grid = lc.sm_count * per_sm;
    if (grid > nrows) grid = nrows;
    const size_t smem = triple_window_smem(win, threads);
    // lock-step batches (grid barrier) pay when H^T is too big to stay in L2 whatever the order of the gathers
    const double ht_bytes = 12.0 * (double)ht_nnz + 4.0 * (double)Ht.rows;
    const bool sync = env_int("SPGEMM_B200_TRIPLE_SYNC", ht_bytes > 40.0e6 ? 1 : 0) != 0 && grid > 1;
    cudaError_t e = cudaErrorUnknown;
    if (sync) {
        e = upper_only ? launch_window<true, true>(grid, threads, smem, lc.stream, H, Q, Ht, row_begin, nrows, win, d_c, d_counters)
                       : launch_window<false, true>(grid, threads, smem, lc.stream, H, Q, Ht, row_begin, nrows, win, d_c, d_counters);
        if (e != cudaSuccess) cudaGetLastError();          // not co-resident (another kernel holds SMs): ticket mode
    }
    if (e != cudaSuccess)
        e = upper_only ? launch_window<true, false>(grid, threads, smem, lc.stream, H, Q, Ht, row_begin, nrows, win, d_c, d_counters)
                       : launch_window<false, false>(grid, threads, smem, lc.stream, H, Q, Ht, row_begin, nrows, win, d_c, d_counters);
    if (e != cudaSuccess) return e;
    ++*lc.launches;
    return cudaSuccess;
}

}  // namespace sb
