// triple.cu -- fused triple product C = H Q H^T into a dense row-major float64 matrix.
//
// Replaces triple_product (/root/reference/src/sparse_sparse_dense.cpp:141-249).  The reference expands
// t = H[i,:] Q into a dense scratch row (:187-198) and then takes a sparse dot of t with EVERY row k >= i of H
// (:201-216, n(n+1)/2 * nnz/row gathers).  Here a thread block owns row i of C: it streams the products
// w = h_ij * q_jc of the expansion (never stored) and contracts each against row c of H^T,
// C[i,k] += w * h_kc for k >= i.  Work is P1 + P2 instead of n^2/2 * nnz/row, H Q is never
// materialised, and every byte of C is written once.
//
// C is built one column panel at a time; a block owns a (panel, row) segment of up to ~26,800 doubles in SHARED memory,
// so every add is a shared-memory add (scripts/micro/atomic_bw.cu: 540 G float64 adds/s chip-wide against 197 G/s for
// L2 reductions), and the slice of H^T a panel gathers from stays in L2.  Two kernels share that structure:
//   k_triple_runs    every row of Q is one run of consecutive columns (banded covariance): a warp streams one
//                    contiguous range of H^T per entry of H and looks the weight up by the column stored with the entry
//   k_triple_panels  any Q: the rows of H^T of 32 products at a time are walked as one flat, balanced stream
// (The round-1 kernel -- a row of C in global memory, every add an L2 reduction, one gather walk per thread -- measured
//  23.2 ms on cfg 5 against 12.4 ms / 21.6 ms for these two: profiles/r2/SUMMARY.md.)
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
// Paneled shared-memory kernel.
//
// C is computed one COLUMN PANEL at a time (TriplePlan: np panels of panel_w columns starting at column k0).  The
// contraction of panel p only needs the entries (k, h_kc) of H^T with k inside the panel -- the p-th part of the
// paneled transpose (analysis.cu), a few tens of MB -- so its gathers hit L2, where the whole H^T (cfg 5: 96 MB
// against ~63 MB of effective L2 for data shared by both dies) made every second gather a DRAM access (round 1:
// 36 GB of DRAM reads, L2 hit rate 50 %, long-scoreboard stalls dominant).  The price is that the expansion
// H[i,:] Q is redone for every panel a row takes part in (upper mode: the panels right of the diagonal), which is
// a coalesced stream of rows of Q.  A panel is at most as wide as the shared-memory accumulator, so every add is a
// shared-memory add and the finished segment leaves with coalesced 128-bit streaming stores.
//
// Work items are (panel, row) pairs handed out panel-major through an atomic ticket.  Per item:
//   1. up to blockDim entries (j, h_ij) of H[i,:] are loaded one per thread together with the extent of row j of Q;
//      a block-wide prefix sum numbers the products of the expansion 0..total-1; every warp takes an equal
//      contiguous share of them, 32 per step;
//   2. software pipeline over the steps of a warp: the loads of (c, q_jc) for step s+2 and of the extent of row c
//      of H^T for step s+1 are in flight while step s is contracted (three dependent memory latencies overlapped);
//   3. contraction as ONE flat stream: the rows of H^T of the warp's 32 products are numbered by a warp prefix sum
//      of their lengths and the warp walks the concatenation 128 entries at a time (four independent loads per
//      lane in flight).  The owner of an entry is found from a 32-bit mask of the row starts inside the step (one
//      warp-wide OR reduction + popcount); its weight w = h_ij q_jc and the offset of its row come from a 32-entry
//      per-warp table in shared memory.  Consecutive products with consecutive c (banded Q) have adjacent rows of
//      H^T, so the stream is coalesced; for any other Q it is a balanced gather.  Every lane is busy whatever the
//      row lengths of H^T (round 1 walked one row per thread: 17.8 of 32 lanes active);
//   4. entries with k >= i (upper mode) are added into the shared segment (float64 CAS loop);
//   5. the segment is streamed out and cleared in one pass.
//
// Dynamic shared memory: acc[win_cap] | hv[nt] | w[nt] | qs[nt] | pre[nt + 1] | d[nt].
struct TripleScratch {
    unsigned red[33];
    unsigned long long cnt[2];
    int item;
};
__host__ __device__ inline size_t triple_window_smem(int win_cap, int threads) {
    return (size_t)win_cap * 8 + (size_t)threads * 28 + 16;
}

struct StepA { int c; double qv, hv; };          // c < 0: no product in this lane
struct StepB { int hs, he; double w; };          // hs == he: nothing to contract

template <bool UPPER>
__global__ void __launch_bounds__(1024, 1)
k_triple_panels(Csr H, Csr Q, const int32_t* __restrict__ t_ptr, const uint32_t* __restrict__ t_pk,
                const double* __restrict__ t_val, TriplePlan plan, int row_begin, int nrows, int win_cap,
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
    unsigned long long p1 = 0, p2_total = 0;
    while (true) {
        if (tid == 0) S.item = (int)atomicAdd(counters + 2, 1ULL);
        __syncthreads();
        int item = S.item;
        __syncthreads();
        // item -> (panel, local row): panel p serves the rows above the end of the panel (upper mode) or all rows
        int p = 0, rows_p = 0;
        for (; p < plan.np; ++p) {
            const int p_end = min(n, plan.k0 + (p + 1) * plan.panel_w);
            rows_p = UPPER ? max(0, min(nrows, p_end - row_begin)) : nrows;
            if (item < rows_p) break;
            item -= rows_p;
        }
        if (p >= plan.np) break;
        const int r = item, i = row_begin + r;
        const int p0 = plan.k0 + p * plan.panel_w, p1c = min(n, p0 + plan.panel_w);
        const int lo = UPPER ? max(i, p0) : p0;                    // first column of the segment
        const bool first_panel = UPPER ? (i >= p0) : (p == 0);     // the panel that holds the diagonal of row i
        double* row = C + (size_t)r * n;
        const int32_t* hp = t_ptr + (size_t)p * H.cols;
        const int h_begin = __ldg(H.ptr + i), h_end = __ldg(H.ptr + i + 1);
        // zeros left of the covered columns: below the diagonal (upper mode) / left of column k0
        if (first_panel) triple_stream_out(row, nullptr, UPPER ? lo : plan.k0);
        unsigned p2 = 0;
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
            __syncthreads();
            if (tid == 0 && first_panel) p1 += total;
            const unsigned wb = (unsigned)((unsigned long long)total * warp / nwarp);
            const unsigned we = (unsigned)((unsigned long long)total * (warp + 1) / nwarp);
            if (wb < we) {
                int r_base;
                {                                          // row of Q of the warp's first product
                    int a = 0, b = cnt;
                    while (b - a > 1) {
                        const int mid = (a + b) >> 1;
                        if (s_pre[mid] <= wb) a = mid; else b = mid;
                    }
                    r_base = a;
                }
                // stage A: which (j, c) is product f, issue the loads of c and q_jc (streamed: used once per row)
                auto stage_a = [&](unsigned f0) {
                    StepA o{-1, 0.0, 0.0};
                    if (f0 >= we) return o;                // warp-uniform
                    const unsigned f = f0 + lane;
                    int rr = r_base;
                    if (f < we) {
                        if (s_pre[rr + 1] <= f) {          // beyond the base row: largest rr with pre[rr] <= f
                            int a = rr + 1, b = cnt;
                            while (b - a > 1) {
                                const int mid = (a + b) >> 1;
                                if (s_pre[mid] <= f) a = mid; else b = mid;
                            }
                            rr = a;
                        }
                        const int q = s_qs[rr] + (int)(f - s_pre[rr]);
                        o.c = __ldcs(Q.idx + q);
                        o.qv = __ldcs(Q.val + q);
                        o.hv = s_hv[rr];
                    }
                    r_base = __shfl_sync(FULL, rr, (int)min(31u, we - f0 - 1u));
                    return o;
                };
                // stage B: weight of the product, issue the loads of the extent of row c of this panel of H^T
                auto stage_b = [&](const StepA& a) {
                    StepB o{0, 0, 0.0};
                    if (a.c >= 0) {
                        o.w = a.hv * a.qv;
                        o.hs = __ldg(hp + a.c);
                        o.he = __ldg(hp + a.c + 1);
                    }
                    return o;
                };
                StepA sa = stage_a(wb);
                StepB sb = stage_b(sa);
                sa = stage_a(wb + 32u);
                for (unsigned f0 = wb; f0 < we; f0 += 32u) {
                    const StepB cur = sb;
                    sb = stage_b(sa);
                    sa = stage_a(f0 + 64u);
                    // stage C: contraction of the 32 products of `cur`
                    const unsigned hlen = (unsigned)(cur.he - cur.hs);
                    const unsigned incl = warp_incl_scan(hlen);
                    const unsigned pre = incl - hlen;
                    const unsigned L = __shfl_sync(FULL, incl, 31);
                    const unsigned nonempty = __ballot_sync(FULL, hlen > 0);
                    if (hlen > 0) {
                        const int rank = __popc(nonempty & lt_mask);
                        s_w[rank] = cur.w;
                        s_d[rank] = cur.hs - (int)pre;
                    }
                    __syncwarp();
                    int heads_before = 0;
                    for (unsigned eb = 0; eb < L; eb += 128u) {
                        int addr[4], k[4];
                        double wv[4], v[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const unsigned e0 = eb + 32u * u;
                            k[u] = -1;
                            addr[u] = 0;
                            wv[u] = 0.0;
                            if (e0 < L) {                  // warp-uniform
                                const unsigned bit = (hlen > 0 && pre >= e0 && pre < e0 + 32u) ? 1u << (pre - e0) : 0u;
                                const unsigned heads = __reduce_or_sync(FULL, bit);
                                const int owner = heads_before + __popc(heads & le_mask) - 1;
                                heads_before += __popc(heads);
                                const unsigned pp = e0 + lane;
                                if (pp < L) {
                                    addr[u] = (int)pp + s_d[owner];
                                    wv[u] = s_w[owner];
                                    k[u] = p0 + (int)(__ldg(t_pk + addr[u]) >> kPanelColBits);
                                }
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) v[u] = k[u] >= lo ? __ldg(t_val + addr[u]) : 0.0;
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            if (k[u] >= lo) {
                                atomicAdd(acc + (k[u] - lo), wv[u] * v[u]);
                                ++p2;
                            }
                        }
                    }
                    __syncwarp();                          // the per-warp tables are rewritten by the next step
                }
            }
            __syncthreads();                               // tables are rewritten by the next slice / item
        }
        p2_total += p2;
        // segment out (coalesced 128-bit streaming stores) and cleared for the next item
        const int count = p1c - lo;
        double* dst = row + lo;
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
        // (the ticket barriers at the top of the loop order the clearing before the next item's adds)
    }
    triple_flush_counters(p1, p2_total, S.cnt, counters);
}

// ---------------------------------------------------------------------------------------------------
// Lean kernel for a Q whose every row is ONE RUN of consecutive ascending columns (a banded covariance -- the
// matrix the triple product is made for; detected by the validation pass, k_check_csr).  Same panels, same
// (panel, row) items and shared-memory segments as k_triple_panels; the difference is the contraction:
//   * the products of one entry h_ij of H are w_t = h_ij * q_{j, c0 + t}, t = 0..len-1, for consecutive columns
//     c0.. of H^T, whose rows are ADJACENT in the panel's CSR: their entries are the one contiguous range
//     [t_ptr[c0], t_ptr[c0 + len]);
//   * a warp takes one entry of H at a time: it puts the (up to 96) weights into its table in shared memory and
//     streams the range with coalesced loads of packed (k, c) words and values (12 bytes per entry); entry
//     (k, c, h_kc) adds w[c - c0] * h_kc into the segment -- the weight is looked up by the low 7 bits of the column
//     stored WITH the entry (a table spans at most 96 consecutive columns), so there is no per-product bookkeeping
//     at all (k_triple_panels spends ~100 instructions per 32 entries on finding owners);
//   * the metadata of all entries of H[i,:] (j, h_ij, extent of row j of Q, c0) is loaded by one thread per entry
//     at the start of the item, so its three dependent gathers are paid once per item, not once per entry, and
//     the weights of a warp's NEXT entry are in flight while it streams the current one.
// Dynamic shared memory: acc[win_cap] | hv[nt] | wt[nwarp * 96] | meta[nt] (int4: start of row j of Q, its length, its
// first column).
constexpr int kRunPiece = 96;                    // columns of a run handled per pass of the weight table
__host__ __device__ inline size_t triple_runs_smem(int win_cap, int threads) {
    return (size_t)win_cap * 8 + (size_t)threads * 48 + 16;
}

// Loads of the panel of H^T carry an L2 evict_last policy: the panel (a few tens of MB) is what every block gathers
// from for the whole pass and should survive the streams of Q and C that pass through L2 beside it.
__device__ __forceinline__ unsigned long long l2_keep_policy() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ int ld_keep_i32(const int32_t* p, unsigned long long pol) {
    int v;
    asm volatile("ld.global.nc.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_keep_u32(const uint32_t* p, unsigned long long pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_keep_f64(const double* p, unsigned long long pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

// (start of row j of Q, its length, its first column) of every entry (i, j) of rows [row_begin, row_begin + nrows) of
// H, once per call: the items of k_triple_runs then read their metadata with one coalesced load instead of three
// dependent gathers per entry, item after item (a row takes part in up to np items).
__global__ void __launch_bounds__(256)
k_triple_entry_meta(Csr H, Csr Q, int row_begin, int nrows, int4* __restrict__ meta) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nrows) return;
    const int r = row_begin + w;
    const int s = __ldg(H.ptr + r), e = __ldg(H.ptr + r + 1);
    for (int p = s + lane_id(); p < e; p += 32) {
        const int j = __ldg(H.idx + p);
        const int qs = __ldg(Q.ptr + j);
        const int len = __ldg(Q.ptr + j + 1) - qs;
        meta[p] = make_int4(qs, len, len > 0 ? __ldg(Q.idx + qs) : 0, 0);
    }
}

constexpr int kRunSlots = 2;                     // independent 32-entry load slots per warp (1: 13.5 ms, 2: 13.3, 4: 14.1)

template <bool UPPER>
__global__ void __launch_bounds__(1024, 1)
k_triple_runs(Csr H, Csr Q, const int32_t* __restrict__ t_ptr, const uint32_t* __restrict__ t_pk,
              const double* __restrict__ t_val, TriplePlan plan, int row_begin, int nrows, int win_cap,
              double* __restrict__ C, unsigned long long* __restrict__ counters,
              const int4* __restrict__ e_meta) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    __shared__ TripleScratch S;
    const int n = H.rows;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    double* acc = reinterpret_cast<double*>(s_raw);
    double* s_hv = acc + win_cap;
    double* s_wt = s_hv + nt + warp * kRunPiece;                    // this warp's 96 weights
    int4* s_meta = reinterpret_cast<int4*>(s_hv + 4 * nt);          // nwarp * 96 == 3 * nt; 16-byte aligned
    const unsigned long long keep = l2_keep_policy();
    if (tid < 2) S.cnt[tid] = 0;
    for (int t = tid; t < win_cap; t += nt) acc[t] = 0.0;
    __syncthreads();
    unsigned long long p1 = 0, p2_total = 0;
    if (tid == 0) S.item = (int)atomicAdd(counters + 2, 1ULL);
    __syncthreads();
    while (true) {
        // the ticket of the NEXT item is drawn while this one is worked on (thread 0 holds it in a register and
        // publishes it at the end of the item), so no item starts with an atomic round trip to L2
        int item = S.item;
        unsigned long long next_ticket = 0;
        __syncthreads();
        if (tid == 0) next_ticket = atomicAdd(counters + 2, 1ULL);
        int p = 0, rows_p = 0;
        for (; p < plan.np; ++p) {
            const int p_end = min(n, plan.k0 + (p + 1) * plan.panel_w);
            rows_p = UPPER ? max(0, min(nrows, p_end - row_begin)) : nrows;
            if (item < rows_p) break;
            item -= rows_p;
        }
        if (p >= plan.np) break;
        const int r = item, i = row_begin + r;
        const int p0 = plan.k0 + p * plan.panel_w, p1c = min(n, p0 + plan.panel_w);
        const int lo = UPPER ? max(i, p0) : p0;
        const bool first_panel = UPPER ? (i >= p0) : (p == 0);
        double* row = C + (size_t)r * n;
        const int32_t* hp = t_ptr + (size_t)p * H.cols;
        const int h_begin = __ldg(H.ptr + i), h_end = __ldg(H.ptr + i + 1);
        if (first_panel) triple_stream_out(row, nullptr, UPPER ? lo : plan.k0);
        unsigned p2 = 0;
        for (int base = h_begin; base < h_end; base += nt) {
            const int cnt = min(nt, h_end - base);
            if (tid < cnt) {                               // one thread per entry of H: its metadata
                int4 m;
                if (e_meta) {
                    m = __ldg(e_meta + base + tid);        // prepared once per call (k_triple_entry_meta)
                } else {
                    const int j = __ldg(H.idx + base + tid);
                    const int qs = __ldg(Q.ptr + j);
                    const int len = __ldg(Q.ptr + j + 1) - qs;
                    m = make_int4(qs, len, len > 0 ? __ldg(Q.idx + qs) : 0, 0);
                }
                s_hv[tid] = __ldg(H.val + base + tid);
                s_meta[tid] = m;
                if (first_panel) p1 += (unsigned)m.y;
            }
            __syncthreads();
            // Software pipeline over this warp's entries of H: while the range of entry e is streamed, the raw
            // values q_{j,c} of entry e + nwarp and the bounds of its range of H^T are already in flight.
            const bool filtered = UPPER && lo > p0;        // the panel holds the diagonal: entries k < i are skipped
            int e = warp;
            double nq0 = 0.0, nq1 = 0.0, nq2 = 0.0, nhv = 0.0;
            int nes = 0, nee = 0;
            int4 nm = make_int4(0, 0, 0, 0);               // metadata of the prefetched entry: one 128-bit broadcast load
            auto prefetch = [&](int en) {
                nq0 = 0.0; nq1 = 0.0; nq2 = 0.0; nes = 0; nee = 0; nhv = 0.0;
                nm = make_int4(0, 0, 0, 0);
                if (en < cnt) {
                    nm = s_meta[en];
                    nhv = s_hv[en];
                    const int nlen = nm.y, nqs = nm.x, nc0 = nm.z;
                    if (lane < nlen) nq0 = __ldcs(Q.val + nqs + lane);
                    if (lane + 32 < nlen) nq1 = __ldcs(Q.val + nqs + lane + 32);
                    if (lane + 64 < nlen) nq2 = __ldcs(Q.val + nqs + lane + 64);
                    if (nlen > 0) {
                        nes = ld_keep_i32(hp + nc0, keep);
                        nee = ld_keep_i32(hp + nc0 + min(kRunPiece, nlen), keep);
                    }
                }
            };
            prefetch(e);
            for (; e < cnt; e += nwarp) {
                const int len = nm.y, qs = nm.x, c0 = nm.z;    // entry e was prefetched: its metadata is in registers
                const double hv = nhv;
                for (int t0 = 0; t0 < len; t0 += kRunPiece) {
                    const int cn = min(kRunPiece, len - t0);
                    const int cb = c0 + t0;
                    double q0, q1, q2;
                    int es, ee;
                    if (t0 == 0) { q0 = nq0; q1 = nq1; q2 = nq2; es = nes; ee = nee; }
                    else {
                        q0 = lane < cn ? __ldcs(Q.val + qs + t0 + lane) : 0.0;
                        q1 = lane + 32 < cn ? __ldcs(Q.val + qs + t0 + lane + 32) : 0.0;
                        q2 = lane + 64 < cn ? __ldcs(Q.val + qs + t0 + lane + 64) : 0.0;
                        es = ld_keep_i32(hp + cb, keep);
                        ee = ld_keep_i32(hp + cb + cn, keep);
                    }
                    if (t0 + kRunPiece >= len) prefetch(e + nwarp);   // last piece of this run: start the next entry's loads
                    s_wt[lane] = hv * q0;
                    s_wt[lane + 32] = hv * q1;
                    s_wt[lane + 64] = hv * q2;
                    __syncwarp();
                    // weight of an entry: s_wt[(c - cb) mod 128], from the low column bits stored with the entry
                    const uint32_t cbm = (uint32_t)cb & kPanelColMask;
                    // The range [es, ee), 32 entries per step (ranges are ~100-200 entries long: wider steps would leave
                    // most lanes of the last one idle), through kRunSlots independent register slots: slot u holds
                    // entries xb + 32 u + lane and is reloaded with its successor (32 kRunSlots entries further) right
                    // before it is added, so the loads of the other slots are in flight whenever the warp waits for
                    // one (no register rotation: a moved register would wait for its load).
                    uint32_t kq[kRunSlots];
                    double vq[kRunSlots];
#pragma unroll
                    for (int u = 0; u < kRunSlots; ++u) {
                        const int xu = es + 32 * u + lane;
                        kq[u] = 0u;
                        vq[u] = 0.0;
                        if (xu < ee) {
                            kq[u] = ld_keep_u32(t_pk + xu, keep);
                            vq[u] = ld_keep_f64(t_val + xu, keep);
                        }
                    }
                    const uint32_t lo_rel = filtered ? (uint32_t)(lo - p0) : 0u;
                    for (int xb = es; xb < ee; xb += 32 * kRunSlots) {
#pragma unroll
                        for (int u = 0; u < kRunSlots; ++u) {
                            const int xu = xb + 32 * u + lane;
                            const uint32_t kc = kq[u];
                            const double v = vq[u];
                            const int xn = xu + 32 * kRunSlots;
                            if (xn < ee) {
                                kq[u] = ld_keep_u32(t_pk + xn, keep);
                                vq[u] = ld_keep_f64(t_val + xn, keep);
                            }
                            const uint32_t krel = kc >> kPanelColBits;
                            if (xu < ee && krel >= lo_rel) {
                                atomicAdd(acc + (krel - lo_rel), s_wt[(kc - cbm) & kPanelColMask] * v);
                                if (filtered) ++p2;
                            }
                        }
                    }
                    if (!filtered && lane == 0) p2 += (unsigned)(ee - es);
                    __syncwarp();                          // the table is rewritten by the next piece
                }
                if (len <= 0) prefetch(e + nwarp);         // empty row of Q: nothing to stream, but the next entry's
                                                           // loads must still be started (its first piece reads them)
            }
            __syncthreads();                               // tables are rewritten by the next slice / item
        }
        p2_total += p2;
        const int count = p1c - lo;
        double* dst = row + lo;
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
        if (tid == 0) S.item = (int)next_ticket;
        __syncthreads();                                   // also orders the clearing before the next item's adds
    }
    triple_flush_counters(p1, p2_total, S.cnt, counters);
}

// ---------------------------------------------------------------------------------------------------
static size_t g_triple_smem_optin = 0, g_triple_smem_sm = 0;

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
    const int dyn = (int)(g_triple_smem_optin - sizeof(TripleScratch) - 64);
    e = cudaFuncSetAttribute(k_triple_panels<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_triple_panels<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_triple_runs<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_triple_runs<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn);
}

static int env_int(const char* name, int dflt) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

static int triple_cap_max() {
    return (int)(((g_triple_smem_optin - sizeof(TripleScratch) - 64 - triple_runs_smem(0, 1024)) / 8) & ~(size_t)1);
}

// Panels of C for rows [row_begin, ...): as few as the shared-memory accumulator allows (a panel is one segment), more
// when the part of H^T they gather from would not stay in L2 (SPGEMM_B200_TRIPLE_L2_MB, default 32 for the 12-byte
// entries and the row pointers of one panel: the data is shared by the SMs of both dies and the streams of Q and C pass
// through beside it; cfg 5 measured 15.1 / 13.9 / 14.0 ms with panels of 32 / 24 / 19 MB); SPGEMM_B200_TRIPLE_PANELS
// forces a count.
TriplePlan triple_plan(int n, int row_begin, bool upper_only, int64_t h_nnz, int h_cols) {
    TriplePlan plan;
    plan.k0 = upper_only ? row_begin : 0;
    const int cover = n - plan.k0 > 0 ? n - plan.k0 : 1;
    const int cap = triple_cap_max();
    int np = (cover + cap - 1) / cap;
    const double covered_bytes = 12.0 * (double)h_nnz * (double)cover / (double)(n > 0 ? n : 1);
    const double budget = 1.0e6 * env_int("SPGEMM_B200_TRIPLE_L2_MB", 32) - 4.0 * (double)h_cols;
    if (budget > 0) {
        const int np_l2 = (int)(covered_bytes / budget) + 1;
        if (np_l2 > np) np = np_l2;
    }
    if (np > 16) np = (cover + cap - 1) / cap > 16 ? (cover + cap - 1) / cap : 16;
    const int forced = env_int("SPGEMM_B200_TRIPLE_PANELS", 0);
    if (forced > 0 && forced >= (cover + cap - 1) / cap) np = forced;
    if (np > cover) np = cover;
    int w = (cover + np - 1) / np;
    w = (w + 1) & ~1;                                   // even width: segments start 16-byte aligned in shared memory
    plan.panel_w = w;
    plan.np = (cover + w - 1) / w;
    return plan;
}

cudaError_t launch_triple_panels(const LaunchCtx& lc, const Csr& H, const Csr& Q, bool q_runs, const int32_t* t_ptr,
                                 const uint32_t* t_pk, const double* t_val, const TriplePlan& plan, bool upper_only,
                                 int row_begin, int nrows, double* d_c, unsigned long long* d_counters,
                                 int4* d_entry_meta) {
    const int n = H.rows;
    if (nrows <= 0 || n <= 0) return cudaSuccess;
    const size_t fixed = sizeof(TripleScratch) + 1024;                    // static scratch + per-block reserve
    const int win = plan.panel_w;
    // 1, 2 or 4 blocks per SM of 1024 / 512 / 256 threads (32 warps per SM at <= 64 registers)
    auto smem_of = [&](int threads) { return q_runs ? triple_runs_smem(win, threads) : triple_window_smem(win, threads); };
    int per_sm = 1;
    for (int cand : {4, 2}) {
        if (g_triple_smem_sm / (smem_of(1024 / cand) + fixed) >= (size_t)cand) { per_sm = cand; break; }
    }
    const int threads = 1024 / per_sm;
    int64_t items = 0;
    for (int p = 0; p < plan.np; ++p) {
        const int p_end = plan.k0 + (p + 1) * plan.panel_w < n ? plan.k0 + (p + 1) * plan.panel_w : n;
        int rows_p = upper_only ? p_end - row_begin : nrows;
        if (rows_p > nrows) rows_p = nrows;
        if (rows_p > 0) items += rows_p;
    }
    int grid = lc.sm_count * per_sm;
    if ((int64_t)grid > items) grid = (int)items;
    if (grid < 1) grid = 1;
    const size_t smem = smem_of(threads);
    if (q_runs) {
        if (d_entry_meta) {
            k_triple_entry_meta<<<(nrows + 7) / 8, 256, 0, lc.stream>>>(H, Q, row_begin, nrows, d_entry_meta);
            SB_LAUNCH_CHECK(lc);
        }
        if (upper_only)
            k_triple_runs<true><<<grid, threads, smem, lc.stream>>>(H, Q, t_ptr, t_pk, t_val, plan, row_begin, nrows, win, d_c, d_counters, d_entry_meta);
        else
            k_triple_runs<false><<<grid, threads, smem, lc.stream>>>(H, Q, t_ptr, t_pk, t_val, plan, row_begin, nrows, win, d_c, d_counters, d_entry_meta);
    } else {
        if (upper_only)
            k_triple_panels<true><<<grid, threads, smem, lc.stream>>>(H, Q, t_ptr, t_pk, t_val, plan, row_begin, nrows, win, d_c, d_counters);
        else
            k_triple_panels<false><<<grid, threads, smem, lc.stream>>>(H, Q, t_ptr, t_pk, t_val, plan, row_begin, nrows, win, d_c, d_counters);
    }
    SB_LAUNCH_CHECK(lc);
    return cudaSuccess;
}

}  // namespace sb
