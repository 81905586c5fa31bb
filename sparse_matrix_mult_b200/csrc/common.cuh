// common.cuh -- shared device helpers for libspgemm_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sb {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kEmpty = -1;          // empty hash slot (column indices are >= 0)

// ---------------------------------------------------------------------------------------------------
// CSR view passed by value to kernels.
struct Csr {
    const int32_t* __restrict__ ptr;
    const int32_t* __restrict__ idx;
    const double* __restrict__ val;
    int rows, cols;
};

// ---------------------------------------------------------------------------------------------------
// warp primitives
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// inclusive warp prefix sum
template <typename T>
__device__ __forceinline__ T warp_incl_scan(T v) {
    const int lane = lane_id();
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T w = __shfl_up_sync(FULL, v, o);
        if (lane >= o) v += w;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread (blockDim.x <= 1024, multiple of 32).
// `scratch` is >= 33 Ts of shared memory.  Returns the exclusive prefix; *total gets the block sum.
template <typename T>
__device__ __forceinline__ T block_excl_scan(T v, T* scratch, T* total) {
    const int lane = lane_id(), warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    T incl = warp_incl_scan(v);
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T s = lane < nwarp ? scratch[lane] : T(0);
        T si = warp_incl_scan(s);
        scratch[lane] = si - s;          // exclusive offset per warp
        if (lane == 31) scratch[32] = si;
    }
    __syncthreads();
    T off = scratch[warp];
    *total = scratch[32];
    __syncthreads();                      // scratch reusable after return
    return off + incl - v;
}

// ---------------------------------------------------------------------------------------------------
// cache-hinted memory operations
// streaming 128-bit store: the dense result is written once and never re-read by the GPU.
__device__ __forceinline__ void st_stream_f64x2(double* p, double a, double b) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ void st_stream_f64(double* p, double a) {
    asm volatile("st.global.cs.f64 [%0], %1;" ::"l"(p), "d"(a) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// hashing: multiplicative hash reduced to [0, size) without a modulo (size need not be a power of two)
__device__ __forceinline__ unsigned hash_slot(int key, unsigned size) {
    return __umulhi(static_cast<unsigned>(key) * 0x9E3779B1u, size);
}

// first position in sorted idx[lo, hi) whose value is >= key
__device__ __forceinline__ int lower_bound(const int32_t* __restrict__ idx, int lo, int hi, int key) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(idx + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---------------------------------------------------------------------------------------------------
// Row expansion: enumerate every intermediate product of one row of A against B.
//
// A warp walks the row's non-zeros 32 at a time: each lane fetches one (j, a_ij) and the extent of row j of
// B, so the 32 dependent gathers (A.idx -> B.ptr) are in flight together.  Rows of B with >= kLongRow
// entries are then streamed by the whole warp (coalesced); the short ones are taken four at a time by
// 8-lane groups, which keeps lanes busy when rows of B have ~10-16 entries.
//
// f(col, prod) is called once per product with prod = a_ij * b_jc (0.0 when WITH_VALUES is false).
// [col_lo, col_hi) restricts the enumeration to a column window: with b_sorted the window bounds are found
// by binary search inside each row of B, otherwise every entry is read and filtered.
// f must not contain warp-synchronous operations (lanes call it divergently).
constexpr int kLongRow = 48;
constexpr int kClipMin = 32;    // rows of B at most this long are filtered, longer ones binary-searched

template <bool WITH_VALUES, class F>
__device__ __forceinline__ void expand_row_warp(const Csr& A, const Csr& B, int a_begin, int a_end,
                                                int col_lo, int col_hi, bool windowed, bool b_sorted, F&& f) {
    const int lane = lane_id();
    const int grp = lane >> 3, gl = lane & 7;
    for (int base = a_begin; base < a_end; base += 32) {
        const int p = base + lane;
        int s = 0, len = 0;
        double av = 0.0;
        if (p < a_end) {
            const int j = __ldg(A.idx + p);
            s = __ldg(B.ptr + j);
            int e = __ldg(B.ptr + j + 1);
            if (WITH_VALUES) av = __ldg(A.val + p);
            if (windowed && b_sorted && e - s > kClipMin) {
                // clip [s, e) to the column window (short rows are cheaper to filter entry by entry)
                if (__ldg(B.idx + s) < col_lo) s = lower_bound(B.idx, s, e, col_lo);
                if (e > s && __ldg(B.idx + e - 1) >= col_hi) e = lower_bound(B.idx, s, e, col_hi);
            }
            len = e - s;
        }
        const bool filter = windowed;
        // whole-warp pass over the long rows of B
        unsigned longmask = __ballot_sync(FULL, len >= kLongRow);
        while (longmask) {
            const int src = __ffs(longmask) - 1;
            longmask &= longmask - 1;
            const int ss = __shfl_sync(FULL, s, src);
            const int ll = __shfl_sync(FULL, len, src);
            const double aa = __shfl_sync(FULL, av, src);
            for (int q = lane; q < ll; q += 32) {
                const int c = __ldg(B.idx + ss + q);
                if (filter && (c < col_lo || c >= col_hi)) continue;
                f(c, WITH_VALUES ? aa * __ldg(B.val + ss + q) : 0.0);
            }
        }
        // 8-lane groups over the short rows
#pragma unroll 1
        for (int e = grp; e < 32; e += 4) {
            const int ss = __shfl_sync(FULL, s, e);
            int ll = __shfl_sync(FULL, len, e);
            const double aa = __shfl_sync(FULL, av, e);
            if (ll >= kLongRow) ll = 0;
            for (int q = gl; q < ll; q += 8) {
                const int c = __ldg(B.idx + ss + q);
                if (filter && (c < col_lo || c >= col_hi)) continue;
                f(c, WITH_VALUES ? aa * __ldg(B.val + ss + q) : 0.0);
            }
        }
    }
}

// Same enumeration by a whole thread block, balanced for power-law inputs (one hub row of B with 10^4 entries
// next to hundreds of rows with 1-8 entries is the normal case there).  The block loads the extents of up to
// blockDim rows of B at once -- one per thread, so the dependent gathers A.idx -> B.ptr are all in flight
// together -- and then
//   * every thread walks its own row of B when that row is short (< kBlockLong entries): no look-up at all;
//   * long rows are cut into slices of kSlice entries; slices are numbered by a block-wide prefix sum and
//     taken by warps round-robin, lanes striding the slice, so the loads of B are coalesced and a hub row is
//     shared by all warps of the block.
// f(col, prod) may be called divergently; it must not contain warp- or block-synchronous operations.
constexpr int kBlockLong = 16;
constexpr int kSlice = 512;

template <int MAXT>
struct SegScratch {
    int start[MAXT];
    int len[MAXT];
    int prefix[MAXT + 1];      // exclusive prefix of the slice counts
    double av[MAXT];
    int red[33];
};

template <bool WITH_VALUES, int MAXT, class F>
__device__ __forceinline__ void expand_row_block(const Csr& A, const Csr& B, int a_begin, int a_end,
                                                 int col_lo, int col_hi, bool windowed, bool b_sorted,
                                                 SegScratch<MAXT>& sc, F&& f) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    const bool filter = windowed;
    for (int base = a_begin; base < a_end; base += nt) {
        const int p = base + tid;
        int s = 0, len = 0;
        double av = 0.0;
        int j = -1;
        if (p < a_end) {
            j = __ldg(A.idx + p);
            if (WITH_VALUES) av = __ldg(A.val + p);
        }
        int e = 0;
        if (j >= 0) {
            s = __ldg(B.ptr + j);
            e = __ldg(B.ptr + j + 1);
        }
        if (j >= 0) {
            if (windowed && b_sorted && e - s > kClipMin) {
                if (__ldg(B.idx + s) < col_lo) s = lower_bound(B.idx, s, e, col_lo);
                if (e > s && __ldg(B.idx + e - 1) >= col_hi) e = lower_bound(B.idx, s, e, col_hi);
            }
            len = e - s;
        }
        const int nslice = len >= kBlockLong ? (len + kSlice - 1) / kSlice : 0;
        int total;
        const int ex = block_excl_scan<int>(nslice, sc.red, &total);
        if (total) {                                   // block-uniform
            sc.start[tid] = s;
            sc.len[tid] = len;
            sc.prefix[tid] = ex;
            if (WITH_VALUES) sc.av[tid] = av;
            if (tid == nt - 1) sc.prefix[nt] = total;
        }
        // short rows: each thread walks its own
        if (len < kBlockLong) {
            for (int q = s; q < s + len; ++q) {
                const int c = __ldg(B.idx + q);
                if (!(filter && (c < col_lo || c >= col_hi))) f(c, WITH_VALUES ? av * __ldg(B.val + q) : 0.0);
            }
        }
        if (total) {
            __syncthreads();
            const int nseg = min(nt, a_end - base);
            for (int item = warp; item < total; item += nwarp) {
                int lo = 0, hi = nseg;                 // prefix[lo] <= item < prefix[hi]; prefix[nseg..nt] == total
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (sc.prefix[mid] <= item) lo = mid; else hi = mid;
                }
                const int off = (item - sc.prefix[lo]) * kSlice;
                const int q0 = sc.start[lo] + off;
                const int cnt = min(kSlice, sc.len[lo] - off);
                const double a = WITH_VALUES ? sc.av[lo] : 0.0;
                // four independent loads per lane in flight before the first use
                const int32_t* bi = B.idx + q0;
                const double* bv = B.val + q0;
                for (int q = lane; q < cnt; q += 128) {
                    int c[4];
                    double v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) c[u] = (q + 32 * u < cnt) ? __ldg(bi + q + 32 * u) : -1;
                    if (WITH_VALUES) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = (q + 32 * u < cnt) ? __ldg(bv + q + 32 * u) : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (c[u] >= 0 && !(filter && (c[u] < col_lo || c[u] >= col_hi)))
                            f(c[u], WITH_VALUES ? a * v[u] : 0.0);
                    }
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace sb
