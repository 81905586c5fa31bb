// api.cu -- the C ABI of libspgemm_b200.so (include/spgemm_b200.h): context, memory, orchestration.
#include <cuda_runtime.h>

#include <emmintrin.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/spgemm_b200.h"
#include "internal.h"

using namespace sb;

// ---------------------------------------------------------------------------------------------------
// handles
struct spgemm_b200_mat {
    int rows, cols;
    int64_t nnz;
    int32_t* ptr;
    int32_t* idx;
    double* val;
    bool owns;
    int32_t* d_sorted;   // device flag, lazily computed (null = unknown): rows sorted by ascending column
    int32_t* d_desc;     // device flag set by transpose: rows sorted by DESCENDING column (null = unknown)
};
struct spgemm_b200_result {
    int rows, cols;
    int64_t nnz;
    int64_t* d_ptr;
    int32_t* d_idx;
    double* d_val;
};

namespace {

enum { EV_START = 0, EV_H2D, EV_ANALYSIS, EV_SYMBOLIC, EV_NUMERIC, EV_POST, EV_D2H, EV_COUNT };

struct Ctx {
    bool ready = false;
    int device = 0;
    int sm_count = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t aux[3] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[3] = {};
    cudaEvent_t ev[EV_COUNT] = {};
    bool ev_pending = false;
    unsigned ev_mask = 0;            // which events were recorded by the current call
    spgemm_b200_stats stats = {};
    int launches = 0;
    void* h_small = nullptr;     // 4 KB pinned staging for counters
    // pinned host cache
    std::mutex host_mu;
    std::multimap<size_t, void*> host_free;
    std::unordered_map<void*, size_t> host_sizes;
    size_t host_cached = 0, host_cache_limit = (size_t)48 << 30;
};
Ctx g;
std::mutex g_mu;
thread_local std::string t_err;

int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
    t_err = what;
    if (e != cudaSuccess) {
        t_err += ": ";
        t_err += cudaGetErrorName(e);
        t_err += " (";
        t_err += cudaGetErrorString(e);
        t_err += ")";
    }
    return code;
}

#define CU(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, #call, e__); \
    } while (0)

int ensure_init() {
    if (g.ready) return SPGEMM_B200_OK;
    const char* env = getenv("SPGEMM_B200_DEVICE");
    return spgemm_b200_init(env ? atoi(env) : 0);
}

LaunchCtx lctx() {
    return LaunchCtx{g.stream, g.sm_count, &g.launches, {g.aux[0], g.aux[1], g.aux[2]}, g.fork_ev,
                     {g.join_ev[0], g.join_ev[1], g.join_ev[2]}};
}

// kernel-variant overrides for experiments: SPGEMM_B200_DENSE_MODE / SPGEMM_B200_TRIPLE_MODE = 0 auto, 1 smem, 2 red
int env_mode(const char* name) {
    const char* v = getenv(name);
    return v ? atoi(v) : 0;
}

template <typename T>
int dalloc(T** p, size_t count) {
    *p = nullptr;
    CU(cudaMallocAsync((void**)p, (count ? count : 1) * sizeof(T), g.stream));
    return SPGEMM_B200_OK;
}
void dfree(void* p) {
    if (p) cudaFreeAsync(p, g.stream);
}

Csr view(const spgemm_b200_mat* m) { return Csr{m->ptr, m->idx, m->val, m->rows, m->cols}; }

int64_t csr_bytes(int64_t rows, int64_t nnz) { return 12 * nnz + 4 * (rows + 1); }

// estimate of products per output element from the operand sizes alone (no pass over the data):
// P ~ nnz(A) * nnz(B) / rows(B)
double products_per_out(const spgemm_b200_mat* a, const spgemm_b200_mat* b) {
    const double out = (double)a->rows * (double)b->cols;
    if (out <= 0 || b->rows <= 0) return 0.0;
    return (double)a->nnz * ((double)b->nnz / (double)b->rows) / out;
}

void mark(int ev) {
    cudaEventRecord(g.ev[ev], g.stream);
    g.ev_mask |= 1u << ev;
}
// `lean` calls (device-resident entry points, which may sit inside a tight timed loop) record only the events
// around their kernels; an unrecorded phase boundary coincides with the previous recorded one.
void begin_call(bool lean = false) {
    g.launches = 0;
    memset(&g.stats, 0, sizeof g.stats);
    g.stats.device = g.device;
    g.ev_mask = 0;
    if (!lean) mark(EV_START);
    g.ev_pending = true;
}

// fold event times into g.stats (blocks until the last recorded event)
void finish_stats() {
    if (!g.ev_pending) return;
    g.ev_pending = false;
    g.stats.launches = g.launches;
    if (!g.ev_mask) return;
    int last = 0, first = EV_COUNT;
    for (int i = 0; i < EV_COUNT; ++i) if (g.ev_mask & (1u << i)) { last = i; if (first == EV_COUNT) first = i; }
    cudaEventSynchronize(g.ev[last]);
    // time of phase ending at event b = elapsed since the closest recorded event before it
    auto phase = [&](int b) {
        if (!(g.ev_mask & (1u << b))) return 0.0;
        int a = b - 1;
        while (a >= 0 && !(g.ev_mask & (1u << a))) --a;
        if (a < 0) return 0.0;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g.ev[a], g.ev[b]);
        return (double)ms;
    };
    g.stats.ms_h2d = phase(EV_H2D);
    g.stats.ms_analysis = phase(EV_ANALYSIS);
    g.stats.ms_symbolic = phase(EV_SYMBOLIC);
    g.stats.ms_numeric = phase(EV_NUMERIC);
    g.stats.ms_post = phase(EV_POST);
    g.stats.ms_d2h += phase(EV_D2H);
    float tot = 0.f;
    if (first != last) cudaEventElapsedTime(&tot, g.ev[first], g.ev[last]);
    g.stats.ms_total += tot;
}

int ensure_sorted_flag(spgemm_b200_mat* m) {
    if (m->d_sorted) return SPGEMM_B200_OK;
    int32_t* buf = nullptr;
    int rc = dalloc(&buf, 4);
    if (rc) return rc;
    CU(launch_check_sorted(lctx(), view(m), m->nnz, buf, buf + 1));
    m->d_sorted = buf;
    return SPGEMM_B200_OK;
}

int check_csr_args(int rows, int cols, const int32_t* ptr, const int32_t* idx, const double* val, const char* name) {
    if (rows < 0 || cols < 0) return fail(SPGEMM_B200_ERR_ARG, name);
    if (!ptr) return fail(SPGEMM_B200_ERR_ARG, name);
    (void)idx; (void)val;
    return SPGEMM_B200_OK;
}

int upload(int rows, int cols, const int32_t* ptr, const int32_t* idx, const double* val, spgemm_b200_mat** out) {
    const int64_t nnz = rows > 0 ? (int64_t)ptr[rows] - ptr[0] : 0;
    if (nnz < 0) return fail(SPGEMM_B200_ERR_ARG, "indptr is not non-decreasing");
    if (nnz > 0 && (!idx || !val)) return fail(SPGEMM_B200_ERR_ARG, "null indices/values with nnz > 0");
    if (rows > 0 && ptr[0] != 0) return fail(SPGEMM_B200_ERR_ARG, "indptr[0] must be 0");
    spgemm_b200_mat* m = new spgemm_b200_mat{rows, cols, nnz, nullptr, nullptr, nullptr, true, nullptr, nullptr};
    int rc;
    if ((rc = dalloc(&m->ptr, (size_t)rows + 1)) || (rc = dalloc(&m->idx, (size_t)nnz)) ||
        (rc = dalloc(&m->val, (size_t)nnz))) {
        spgemm_b200_mat_free(m);
        return rc;
    }
    cudaError_t e = cudaSuccess;
    if (rows > 0) e = cudaMemcpyAsync(m->ptr, ptr, ((size_t)rows + 1) * 4, cudaMemcpyHostToDevice, g.stream);
    else e = cudaMemsetAsync(m->ptr, 0, 4, g.stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->idx, idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, g.stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->val, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, g.stream);
    if (e != cudaSuccess) {
        spgemm_b200_mat_free(m);
        return fail(SPGEMM_B200_ERR_CUDA, "operand upload", e);
    }
    g.stats.bytes_h2d += csr_bytes(rows, nnz);
    *out = m;
    return SPGEMM_B200_OK;
}

// Build X^T on the device (rows of the transpose are in arbitrary order).
int transpose_impl(const spgemm_b200_mat* x, spgemm_b200_mat** out) {
    spgemm_b200_mat* t = new spgemm_b200_mat{x->cols, x->rows, x->nnz, nullptr, nullptr, nullptr, true, nullptr, nullptr};
    int32_t *counts = nullptr, *cursor = nullptr;
    int64_t* tmp = nullptr;
    int rc;
    if ((rc = dalloc(&t->ptr, (size_t)t->rows + 1)) || (rc = dalloc(&t->idx, (size_t)t->nnz)) ||
        (rc = dalloc(&t->val, (size_t)t->nnz)) || (rc = dalloc(&counts, (size_t)t->rows + 1)) ||
        (rc = dalloc(&cursor, (size_t)t->rows + 1)) || (rc = dalloc(&tmp, 1032))) {
        spgemm_b200_mat_free(t); dfree(counts); dfree(cursor); dfree(tmp);
        return rc;
    }
    LaunchCtx lc = lctx();
    cudaError_t e = cudaMemsetAsync(counts, 0, ((size_t)t->rows + 1) * 4, g.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(cursor, 0, ((size_t)t->rows + 1) * 4, g.stream);
    if (e == cudaSuccess) e = launch_transpose_count(lc, view(x), x->nnz, counts);
    if (e == cudaSuccess) e = launch_scan_i32(lc, counts, t->ptr, t->rows, tmp);
    if (e == cudaSuccess) e = launch_transpose_fill(lc, view(x), t->ptr, cursor, t->idx, t->val);
    // rows of the transpose come out in atomic order: sort them by descending column (see k_sort_rows_desc)
    if (e == cudaSuccess && dalloc(&t->d_desc, 1) == SPGEMM_B200_OK) {
        const int32_t one = 1;
        e = cudaMemcpyAsync(t->d_desc, &one, 4, cudaMemcpyHostToDevice, g.stream);
        if (e == cudaSuccess) e = launch_sort_rows_desc(lc, t->rows, t->ptr, t->idx, t->val, t->d_desc);
    }
    dfree(counts); dfree(cursor); dfree(tmp);
    if (e != cudaSuccess) {
        spgemm_b200_mat_free(t);
        return fail(SPGEMM_B200_ERR_CUDA, "transpose", e);
    }
    *out = t;
    return SPGEMM_B200_OK;
}

// Zero `count` doubles with non-temporal stores: the destination is not read first (memset below its internal
// threshold write-allocates, doubling the memory traffic) and the zeros do not displace the cache.
static void zero_nt(double* p, size_t count) {
    while (count && (reinterpret_cast<uintptr_t>(p) & 15)) { *p++ = 0.0; --count; }
    __m128i z = _mm_setzero_si128();
    __m128i* v = reinterpret_cast<__m128i*>(p);
    size_t nv = count / 2;
    size_t i = 0;
    for (; i + 4 <= nv; i += 4) {
        _mm_stream_si128(v + i, z);
        _mm_stream_si128(v + i + 1, z);
        _mm_stream_si128(v + i + 2, z);
        _mm_stream_si128(v + i + 3, z);
    }
    for (; i < nv; ++i) _mm_stream_si128(v + i, z);
    if (count & 1) p[count - 1] = 0.0;
}

// Device -> host copy of an n x n result whose strictly lower triangle is known to be zero (symmetric dense mode,
// triple product): only the upper trapezoids cross PCIe -- row block [r0, r1) sends columns [r0, n) as one 2-D
// copy -- while host threads zero the rectangles to their left.  Halves the bytes on the link, which is what
// bounds these modes end to end (3.2 GB at ~55 GB/s for BASELINE config 2).
cudaError_t d2h_upper(const double* d_c, int n, double* c_host) {
    if (n <= 0) return cudaSuccess;
    {
        const int nb = n < 64 ? 1 : 64, st = (n + nb - 1) / nb;
        for (int r0 = 0; r0 < n; r0 += st) g.stats.bytes_d2h += (int64_t)(n - r0) * ((r0 + st < n ? r0 + st : n) - r0) * 8;
    }
    const int blocks = n < 64 ? 1 : 64;
    const int step = (n + blocks - 1) / blocks;
    cudaError_t err = cudaSuccess;
    for (int r0 = 0; r0 < n && err == cudaSuccess; r0 += step) {
        const int r1 = r0 + step < n ? r0 + step : n;
        err = cudaMemcpy2DAsync(c_host + (size_t)r0 * n + r0, (size_t)n * 8, d_c + (size_t)r0 * n + r0, (size_t)n * 8,
                                (size_t)(n - r0) * 8, (size_t)(r1 - r0), cudaMemcpyDeviceToHost, g.stream);
    }
    // zero the lower-left rectangles on the host meanwhile (rows are split evenly by AREA over the threads)
    unsigned hw = std::thread::hardware_concurrency();
    int nthreads = hw == 0 ? 4 : (hw > 8 ? 8 : (int)hw);     // measured: flat beyond 4-8 threads (~40 GB/s)
    if (const char* ev = getenv("SPGEMM_B200_ZERO_THREADS")) nthreads = atoi(ev);      // 0 = skip (experiments only)
    if ((size_t)n * n < ((size_t)1 << 22) && nthreads > 1) nthreads = 1;
    auto zero_rows = [=](int t) {
        // thread t takes row blocks t, t + nthreads, ... (interleaved: equal area per thread)
        int b = 0;
        for (int r0 = 0; r0 < n; r0 += step, ++b) {
            if (b % nthreads != t || r0 == 0) continue;
            const int r1 = r0 + step < n ? r0 + step : n;
            for (int r = r0; r < r1; ++r) zero_nt(c_host + (size_t)r * n, (size_t)r0);
        }
        _mm_sfence();
    };
    if (nthreads <= 0) return err;
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(zero_rows, t);
    zero_rows(0);
    for (auto& th : pool) th.join();
    return err;
}

// Sparse product of rows [r0, r1).  Records EV_ANALYSIS / EV_SYMBOLIC / EV_NUMERIC.
int csr_impl(spgemm_b200_mat* a, spgemm_b200_mat* b, int upper_only, int r0, int r1, spgemm_b200_result** out) {
    const int m = r1 - r0, n = b->cols;
    spgemm_b200_result* res = new spgemm_b200_result{m, n, 0, nullptr, nullptr, nullptr};
    int rc = dalloc(&res->d_ptr, (size_t)m + 1);
    if (rc) { delete res; return rc; }
    g.stats.bytes_min = csr_bytes(m, (int64_t)0) + csr_bytes(b->rows, b->nnz);
    if (m == 0 || a->nnz == 0 || b->nnz == 0) {
        CU(cudaMemsetAsync(res->d_ptr, 0, ((size_t)m + 1) * 8, g.stream));
        mark(EV_ANALYSIS); mark(EV_SYMBOLIC); mark(EV_NUMERIC);
        rc = dalloc(&res->d_idx, 1);
        if (!rc) rc = dalloc(&res->d_val, 1);
        if (rc) { spgemm_b200_result_free(res); return rc; }
        *out = res;
        return SPGEMM_B200_OK;
    }
    if ((rc = ensure_sorted_flag(b))) { spgemm_b200_result_free(res); return rc; }

    // one workspace block: nnz[m] | lists[BINS*m] | small counters | scan scratch
    const int bins = SYM_BINS > NUM_BINS ? SYM_BINS : NUM_BINS;
    const size_t small_ints = 32;    // cursor[16] | work[8] | pad ; total (u64) lives at small + 24
    const size_t ws_ints = (size_t)m + (size_t)bins * m + small_ints;
    int32_t* ws = nullptr;
    int64_t* scan_tmp = nullptr;
    if ((rc = dalloc(&ws, ws_ints)) || (rc = dalloc(&scan_tmp, 1032))) {
        dfree(ws); spgemm_b200_result_free(res);
        return rc;
    }
    int32_t* d_nnz = ws;
    int32_t* d_lists = ws + m;
    int32_t* d_small = d_lists + (size_t)bins * m;
    // keep the 8-byte total aligned: d_small offset must be even
    if ((reinterpret_cast<uintptr_t>(d_small) & 7) != 0) d_small += 1;   // ws_ints has slack (32 > 16+8+2+1)
    int32_t* d_cursor = d_small;
    int32_t* d_work = d_small + 16;
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(d_small + 24);
    auto bail = [&](int code) {
        dfree(ws); dfree(scan_tmp); spgemm_b200_result_free(res);
        return code;
    };
    LaunchCtx lc = lctx();
    SparseJob job{view(a), view(b), r0, m, upper_only != 0, b->d_sorted};
    cudaError_t e = cudaMemsetAsync(d_small, 0, 28 * sizeof(int32_t), g.stream);
    if (e == cudaSuccess)
        e = launch_row_products(lc, job.A, job.B, r0, m, job.upper_only, nullptr, d_nnz, d_lists, d_cursor, d_total);
    if (e != cudaSuccess) return bail(fail(SPGEMM_B200_ERR_CUDA, "row products", e));
    mark(EV_ANALYSIS);
    int32_t* h = static_cast<int32_t*>(g.h_small);
    e = cudaMemcpyAsync(h, d_small, 28 * sizeof(int32_t), cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) return bail(fail(SPGEMM_B200_ERR_CUDA, "bin counts", e));
    int32_t sym_counts[SYM_BINS];
    for (int k = 0; k < SYM_BINS; ++k) sym_counts[k] = h[k];
    unsigned long long total_products;
    memcpy(&total_products, h + 24, 8);
    g.stats.products = (int64_t)total_products;

    e = launch_symbolic(lc, job, d_lists, sym_counts, d_nnz, d_work);
    if (e == cudaSuccess) e = launch_scan_i64(lc, d_nnz, res->d_ptr, m, scan_tmp);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_cursor, 0, 16 * sizeof(int32_t), g.stream);
    if (e == cudaSuccess) e = launch_bin_by_nnz(lc, d_nnz, m, d_lists, d_cursor);
    if (e != cudaSuccess) return bail(fail(SPGEMM_B200_ERR_CUDA, "symbolic phase", e));
    mark(EV_SYMBOLIC);
    int64_t* h64 = reinterpret_cast<int64_t*>(h + 32);
    e = cudaMemcpyAsync(h, d_cursor, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h64, res->d_ptr + m, 8, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) return bail(fail(SPGEMM_B200_ERR_CUDA, "nnz(C)", e));
    int32_t num_counts[NUM_BINS];
    for (int k = 0; k < NUM_BINS; ++k) num_counts[k] = h[k];
    res->nnz = *h64;
    g.stats.nnz_c = res->nnz;
    // bytes(A) + bytes(B) + bytes(C); a row slice of A is charged pro rata
    g.stats.bytes_min = csr_bytes(m, a->rows ? a->nnz * m / a->rows : 0) + csr_bytes(b->rows, b->nnz) + csr_bytes(m, res->nnz);
    if ((rc = dalloc(&res->d_idx, (size_t)res->nnz)) || (rc = dalloc(&res->d_val, (size_t)res->nnz))) return bail(rc);
    e = launch_numeric(lc, job, d_lists, num_counts, res->d_ptr, res->d_idx, res->d_val, d_work);
    if (e != cudaSuccess) return bail(fail(SPGEMM_B200_ERR_CUDA, "numeric phase", e));
    mark(EV_NUMERIC);
    dfree(ws); dfree(scan_tmp);
    *out = res;
    return SPGEMM_B200_OK;
}

}  // namespace

// ===================================================================================================
extern "C" {

const char* spgemm_b200_version(void) { return "spgemm_b200 0.1 (sm_100a)"; }

const char* spgemm_b200_last_error(void) { return t_err.c_str(); }

int spgemm_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int spgemm_b200_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.ready && g.device == device) return SPGEMM_B200_OK;
    if (g.ready) return fail(SPGEMM_B200_ERR_STATE, "already initialised on another device; call spgemm_b200_shutdown first");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SPGEMM_B200_ERR_CUDA, "no CUDA device available (libspgemm_b200 has no CPU fallback)", e);
    }
    if (device < 0 || device >= n) return fail(SPGEMM_B200_ERR_ARG, "device ordinal out of range");
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail(SPGEMM_B200_ERR_CUDA, "device is not sm_100 class; this library carries sm_100a code only");
    g.device = device;
    g.sm_count = prop.multiProcessorCount;
    CU(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    for (int i = 0; i < EV_COUNT; ++i) CU(cudaEventCreate(&g.ev[i]));
    for (int i = 0; i < 3; ++i) {
        CU(cudaStreamCreateWithFlags(&g.aux[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&g.join_ev[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&g.fork_ev, cudaEventDisableTiming));
    CU(cudaHostAlloc(&g.h_small, 4096, cudaHostAllocDefault));
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;       // cache freed blocks: repeated calls reuse their workspaces
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    CU(sparse_kernels_configure());
    CU(dense_kernels_configure());
    CU(triple_kernels_configure());
    const char* lim = getenv("SPGEMM_B200_PINNED_CACHE_GB");
    if (lim) g.host_cache_limit = (size_t)atoll(lim) << 30;
    g.ready = true;
    return SPGEMM_B200_OK;
}

void spgemm_b200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g.ready) return;
    cudaStreamSynchronize(g.own_stream);
    {
        std::lock_guard<std::mutex> hl(g.host_mu);
        for (auto& kv : g.host_free) cudaFreeHost(kv.second);
        g.host_free.clear();
        g.host_cached = 0;
    }
    for (int i = 0; i < EV_COUNT; ++i) cudaEventDestroy(g.ev[i]);
    for (int i = 0; i < 3; ++i) { cudaStreamDestroy(g.aux[i]); cudaEventDestroy(g.join_ev[i]); g.aux[i] = nullptr; }
    cudaEventDestroy(g.fork_ev);
    cudaFreeHost(g.h_small);
    cudaStreamDestroy(g.own_stream);
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, g.device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
    g.own_stream = g.stream = nullptr;
    g.ready = false;
}

int spgemm_b200_get_stats(spgemm_b200_stats* out) {
    if (!out) return fail(SPGEMM_B200_ERR_ARG, "null stats");
    if (!g.ready) return fail(SPGEMM_B200_ERR_STATE, "not initialised");
    finish_stats();
    *out = g.stats;
    return SPGEMM_B200_OK;
}

int spgemm_b200_set_stream(void* stream) {
    int rc = ensure_init();
    if (rc) return rc;
    g.stream = stream ? static_cast<cudaStream_t>(stream) : g.own_stream;
    return SPGEMM_B200_OK;
}

int spgemm_b200_synchronize(void) {
    int rc = ensure_init();
    if (rc) return rc;
    CU(cudaStreamSynchronize(g.stream));
    return SPGEMM_B200_OK;
}

// ---- pinned host cache ------------------------------------------------------------------------------
void* spgemm_b200_host_alloc(size_t bytes) {
    if (ensure_init()) return nullptr;
    const size_t granule = (size_t)1 << 20;
    const size_t size = ((bytes ? bytes : 1) + granule - 1) / granule * granule;
    {
        std::lock_guard<std::mutex> hl(g.host_mu);
        // smallest cached buffer that fits, as long as it wastes less than half of itself
        auto it = g.host_free.lower_bound(size);
        if (it != g.host_free.end() && it->first <= 2 * size) {
            void* p = it->second;
            g.host_cached -= it->first;
            g.host_free.erase(it);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, size, cudaHostAllocDefault);
    if (e != cudaSuccess) {
        // drop the cache and retry once
        {
            std::lock_guard<std::mutex> hl(g.host_mu);
            for (auto& kv : g.host_free) { cudaFreeHost(kv.second); g.host_sizes.erase(kv.second); }
            g.host_free.clear();
            g.host_cached = 0;
        }
        cudaGetLastError();
        e = cudaHostAlloc(&p, size, cudaHostAllocDefault);
        if (e != cudaSuccess) { fail(SPGEMM_B200_ERR_CUDA, "cudaHostAlloc", e); return nullptr; }
    }
    std::lock_guard<std::mutex> hl(g.host_mu);
    g.host_sizes[p] = size;
    return p;
}

void spgemm_b200_host_free(void* p) {
    if (!p || !g.ready) return;
    std::lock_guard<std::mutex> hl(g.host_mu);
    auto it = g.host_sizes.find(p);
    if (it == g.host_sizes.end()) return;
    const size_t size = it->second;
    if (g.host_cached + size <= g.host_cache_limit) {
        g.host_free.emplace(size, p);
        g.host_cached += size;
    } else {
        g.host_sizes.erase(it);
        cudaFreeHost(p);
    }
}

// ---- matrices -----------------------------------------------------------------------------------------
int spgemm_b200_mat_upload(int rows, int cols, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                           const double* values, spgemm_b200_mat** out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!out) return fail(SPGEMM_B200_ERR_ARG, "null out");
    if ((rc = check_csr_args(rows, cols, indptr, indices, values, "mat_upload: bad matrix"))) return rc;
    if (rows > 0 && (int64_t)indptr[rows] != nnz) return fail(SPGEMM_B200_ERR_ARG, "mat_upload: nnz != indptr[rows]");
    return upload(rows, cols, indptr, indices, values, out);
}

int spgemm_b200_mat_wrap(int rows, int cols, int64_t nnz, const int32_t* d_indptr, const int32_t* d_indices,
                         const double* d_values, spgemm_b200_mat** out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!out || !d_indptr || rows < 0 || cols < 0 || nnz < 0) return fail(SPGEMM_B200_ERR_ARG, "mat_wrap: bad argument");
    *out = new spgemm_b200_mat{rows, cols, nnz, const_cast<int32_t*>(d_indptr), const_cast<int32_t*>(d_indices),
                               const_cast<double*>(d_values), false, nullptr, nullptr};
    return SPGEMM_B200_OK;
}

int spgemm_b200_mat_transpose(const spgemm_b200_mat* x, spgemm_b200_mat** out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!x || !out) return fail(SPGEMM_B200_ERR_ARG, "mat_transpose: null argument");
    return transpose_impl(x, out);
}

void spgemm_b200_mat_free(spgemm_b200_mat* m) {
    if (!m) return;
    if (m->owns) { dfree(m->ptr); dfree(m->idx); dfree(m->val); }
    dfree(m->d_sorted);
    dfree(m->d_desc);
    delete m;
}

// ---- sparse output --------------------------------------------------------------------------------------
int spgemm_b200_csr_dev(const spgemm_b200_mat* a, const spgemm_b200_mat* b, int upper_only, int row_begin, int row_end,
                        spgemm_b200_result** out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!a || !b || !out) return fail(SPGEMM_B200_ERR_ARG, "csr_dev: null argument");
    if (a->cols != b->rows) return fail(SPGEMM_B200_ERR_ARG, "csr_dev: inner dimensions differ");
    if (row_end < 0) { row_begin = 0; row_end = a->rows; }
    if (row_begin < 0 || row_end > a->rows || row_begin > row_end) return fail(SPGEMM_B200_ERR_ARG, "csr_dev: bad row range");
    begin_call();
    mark(EV_H2D);
    rc = csr_impl(const_cast<spgemm_b200_mat*>(a), const_cast<spgemm_b200_mat*>(b), upper_only, row_begin, row_end, out);
    mark(EV_POST); mark(EV_D2H);
    return rc;
}

int spgemm_b200_csr(int m, int k, int n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_values,
                    const int32_t* b_indptr, const int32_t* b_indices, const double* b_values, int upper_only,
                    spgemm_b200_result** out) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!out) return fail(SPGEMM_B200_ERR_ARG, "csr: null out");
    if ((rc = check_csr_args(m, k, a_indptr, a_indices, a_values, "csr: bad A"))) return rc;
    if ((rc = check_csr_args(k, n, b_indptr, b_indices, b_values, "csr: bad B"))) return rc;
    begin_call();
    spgemm_b200_mat *a = nullptr, *b = nullptr;
    if ((rc = upload(m, k, a_indptr, a_indices, a_values, &a))) return rc;
    const bool same = (a_indptr == b_indptr && a_indices == b_indices && a_values == b_values && m == k && k == n);
    if (same) b = a;
    else if ((rc = upload(k, n, b_indptr, b_indices, b_values, &b))) { spgemm_b200_mat_free(a); return rc; }
    mark(EV_H2D);
    rc = csr_impl(a, b, upper_only, 0, m, out);
    mark(EV_POST); mark(EV_D2H);
    spgemm_b200_mat_free(a);
    if (!same) spgemm_b200_mat_free(b);
    return rc;
}

int64_t spgemm_b200_result_nnz(const spgemm_b200_result* r) { return r ? r->nnz : -1; }
int spgemm_b200_result_rows(const spgemm_b200_result* r) { return r ? r->rows : -1; }
int spgemm_b200_result_cols(const spgemm_b200_result* r) { return r ? r->cols : -1; }

int spgemm_b200_result_device_ptrs(const spgemm_b200_result* r, const int64_t** d_indptr, const int32_t** d_indices,
                                   const double** d_values) {
    if (!r) return fail(SPGEMM_B200_ERR_ARG, "null result");
    if (d_indptr) *d_indptr = r->d_ptr;
    if (d_indices) *d_indices = r->d_idx;
    if (d_values) *d_values = r->d_val;
    return SPGEMM_B200_OK;
}

int spgemm_b200_result_copy(const spgemm_b200_result* r, void* indptr, int index64, int32_t* indices, double* values) {
    if (!g.ready) return fail(SPGEMM_B200_ERR_STATE, "not initialised");
    if (!r || !indptr) return fail(SPGEMM_B200_ERR_ARG, "result_copy: null argument");
    if (r->nnz > 0 && (!indices || !values)) return fail(SPGEMM_B200_ERR_ARG, "result_copy: null indices/values");
    if (!index64 && r->nnz > 0x7fffffffLL) return fail(SPGEMM_B200_ERR_OVERFLOW, "nnz(C) >= 2^31 needs index64");
    finish_stats();
    cudaEvent_t e0 = g.ev[EV_POST], e1 = g.ev[EV_D2H];
    CU(cudaEventRecord(e0, g.stream));
    int32_t* narrow = nullptr;
    if (index64) {
        CU(cudaMemcpyAsync(indptr, r->d_ptr, ((size_t)r->rows + 1) * 8, cudaMemcpyDeviceToHost, g.stream));
    } else {
        int rc = dalloc(&narrow, (size_t)r->rows + 1);
        if (rc) return rc;
        CU(launch_narrow_indptr(lctx(), r->d_ptr, narrow, r->rows + 1));
        CU(cudaMemcpyAsync(indptr, narrow, ((size_t)r->rows + 1) * 4, cudaMemcpyDeviceToHost, g.stream));
    }
    if (r->nnz > 0) {
        CU(cudaMemcpyAsync(indices, r->d_idx, (size_t)r->nnz * 4, cudaMemcpyDeviceToHost, g.stream));
        CU(cudaMemcpyAsync(values, r->d_val, (size_t)r->nnz * 8, cudaMemcpyDeviceToHost, g.stream));
    }
    CU(cudaEventRecord(e1, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    dfree(narrow);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    g.stats.bytes_d2h += (int64_t)(r->rows + 1) * (index64 ? 8 : 4) + r->nnz * 12;
    g.stats.ms_d2h += ms;
    g.stats.ms_total += ms;
    g.stats.launches = g.launches;
    return SPGEMM_B200_OK;
}

void spgemm_b200_result_free(spgemm_b200_result* r) {
    if (!r) return;
    dfree(r->d_ptr); dfree(r->d_idx); dfree(r->d_val);
    delete r;
}

// ---- dense output ---------------------------------------------------------------------------------------
int spgemm_b200_dense_dev(const spgemm_b200_mat* a, const spgemm_b200_mat* b, int upper_only, int row_begin, int row_end,
                          double* d_c) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!a || !b || !d_c) return fail(SPGEMM_B200_ERR_ARG, "dense_dev: null argument");
    if (a->cols != b->rows) return fail(SPGEMM_B200_ERR_ARG, "dense_dev: inner dimensions differ");
    if (row_end < 0) { row_begin = 0; row_end = a->rows; }
    if (row_begin < 0 || row_end > a->rows || row_begin > row_end) return fail(SPGEMM_B200_ERR_ARG, "dense_dev: bad row range");
    begin_call(true);
    if (!b->d_sorted) {
        mark(EV_H2D);
        if ((rc = ensure_sorted_flag(const_cast<spgemm_b200_mat*>(b)))) return rc;
    }
    mark(EV_SYMBOLIC);
    cudaError_t de = launch_dense(lctx(), view(a), view(b), b->d_sorted, upper_only != 0, row_begin, row_end - row_begin,
                                  d_c, env_mode("SPGEMM_B200_DENSE_MODE"), products_per_out(a, b));
    if (de != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "dense kernel", de);
    mark(EV_NUMERIC);
    g.stats.nnz_c = (int64_t)(row_end - row_begin) * b->cols;
    g.stats.bytes_min = csr_bytes(a->rows, a->nnz) + csr_bytes(b->rows, b->nnz) + 8 * g.stats.nnz_c;
    return SPGEMM_B200_OK;
}

int spgemm_b200_dense(int m, int k, int n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_values,
                      const int32_t* b_indptr, const int32_t* b_indices, const double* b_values, int upper_only,
                      int mirror, double* c_host) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!c_host && (int64_t)m * n > 0) return fail(SPGEMM_B200_ERR_ARG, "dense: null output");
    if ((rc = check_csr_args(m, k, a_indptr, a_indices, a_values, "dense: bad A"))) return rc;
    if ((rc = check_csr_args(k, n, b_indptr, b_indices, b_values, "dense: bad B"))) return rc;
    if (mirror && (!upper_only || m != n)) return fail(SPGEMM_B200_ERR_ARG, "dense: mirror needs upper_only and a square result");
    begin_call();
    spgemm_b200_mat *a = nullptr, *b = nullptr;
    if ((rc = upload(m, k, a_indptr, a_indices, a_values, &a))) return rc;
    if ((rc = upload(k, n, b_indptr, b_indices, b_values, &b))) { spgemm_b200_mat_free(a); return rc; }
    mark(EV_H2D);
    double* d_c = nullptr;
    auto done = [&](int code) {
        dfree(d_c); spgemm_b200_mat_free(a); spgemm_b200_mat_free(b);
        return code;
    };
    if ((rc = ensure_sorted_flag(b))) return done(rc);
    mark(EV_ANALYSIS); mark(EV_SYMBOLIC);
    const size_t elems = (size_t)m * (size_t)n;
    if ((rc = dalloc(&d_c, elems))) return done(rc);
    cudaError_t e = launch_dense(lctx(), view(a), view(b), b->d_sorted, upper_only != 0, 0, m, d_c,
                                 env_mode("SPGEMM_B200_DENSE_MODE"), products_per_out(a, b));
    if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "dense kernel", e));
    mark(EV_NUMERIC);
    if (mirror) {
        e = launch_mirror(lctx(), d_c, n);
        if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "mirror kernel", e));
    }
    mark(EV_POST);
    if (elems) {
        if (upper_only && !mirror && m == n) e = d2h_upper(d_c, n, c_host);
        else { e = cudaMemcpyAsync(c_host, d_c, elems * 8, cudaMemcpyDeviceToHost, g.stream); g.stats.bytes_d2h += (int64_t)elems * 8; }
        if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "dense result copy", e));
    }
    mark(EV_D2H);
    e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "dense synchronize", e));
    g.stats.nnz_c = (int64_t)elems;
    g.stats.bytes_min = csr_bytes(m, a->nnz) + csr_bytes(k, b->nnz) + 8 * (int64_t)elems * (mirror ? 2 : 1);
    return done(SPGEMM_B200_OK);
}

// ---- triple product -------------------------------------------------------------------------------------
int spgemm_b200_triple_dev(const spgemm_b200_mat* h, const spgemm_b200_mat* q, const spgemm_b200_mat* ht, int upper_only,
                           int row_begin, int row_end, double* d_c) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!h || !q || !d_c) return fail(SPGEMM_B200_ERR_ARG, "triple_dev: null argument");
    if (h->cols != q->rows || q->rows != q->cols) return fail(SPGEMM_B200_ERR_ARG, "triple_dev: Q must be square with H.cols rows");
    if (row_end < 0) { row_begin = 0; row_end = h->rows; }
    if (row_begin < 0 || row_end > h->rows || row_begin > row_end) return fail(SPGEMM_B200_ERR_ARG, "triple_dev: bad row range");
    begin_call();
    mark(EV_H2D);
    spgemm_b200_mat* own_ht = nullptr;
    if (!ht) {
        if ((rc = transpose_impl(h, &own_ht))) return rc;
        ht = own_ht;
    }
    mark(EV_ANALYSIS); mark(EV_SYMBOLIC);
    unsigned long long* d_cnt = nullptr;
    if ((rc = dalloc(&d_cnt, 4))) { spgemm_b200_mat_free(own_ht); return rc; }
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, 32, g.stream);
    if (e == cudaSuccess)
        e = launch_triple(lctx(), view(h), view(q), view(ht), ht->d_desc, upper_only != 0, row_begin, row_end - row_begin, d_c, d_cnt,
                          env_mode("SPGEMM_B200_TRIPLE_MODE"));
    mark(EV_NUMERIC); mark(EV_POST);
    unsigned long long* hc = reinterpret_cast<unsigned long long*>(static_cast<char*>(g.h_small) + 512);
    if (e == cudaSuccess) e = cudaMemcpyAsync(hc, d_cnt, 16, cudaMemcpyDeviceToHost, g.stream);
    mark(EV_D2H);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    dfree(d_cnt);
    spgemm_b200_mat_free(own_ht);
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "triple kernel", e);
    g.stats.products = (int64_t)(hc[0] + hc[1]);
    g.stats.nnz_c = (int64_t)(row_end - row_begin) * h->rows;
    g.stats.bytes_min = 2 * csr_bytes(h->rows, h->nnz) + csr_bytes(q->rows, q->nnz) + 8 * g.stats.nnz_c;
    return SPGEMM_B200_OK;
}

int spgemm_b200_triple(int n, int k, const int32_t* h_indptr, const int32_t* h_indices, const double* h_values,
                       const int32_t* q_indptr, const int32_t* q_indices, const double* q_values, int mode,
                       double* c_host) {
    int rc = ensure_init();
    if (rc) return rc;
    if (mode < 0 || mode > 2) return fail(SPGEMM_B200_ERR_ARG, "triple: bad mode");
    if (!c_host && n > 0) return fail(SPGEMM_B200_ERR_ARG, "triple: null output");
    if ((rc = check_csr_args(n, k, h_indptr, h_indices, h_values, "triple: bad H"))) return rc;
    if ((rc = check_csr_args(k, k, q_indptr, q_indices, q_values, "triple: bad Q"))) return rc;
    begin_call();
    spgemm_b200_mat *h = nullptr, *q = nullptr, *ht = nullptr;
    double* d_c = nullptr;
    unsigned long long* d_cnt = nullptr;
    auto done = [&](int code) {
        dfree(d_c); dfree(d_cnt);
        spgemm_b200_mat_free(h); spgemm_b200_mat_free(q); spgemm_b200_mat_free(ht);
        return code;
    };
    if ((rc = upload(n, k, h_indptr, h_indices, h_values, &h))) return done(rc);
    if ((rc = upload(k, k, q_indptr, q_indices, q_values, &q))) return done(rc);
    mark(EV_H2D);
    if ((rc = transpose_impl(h, &ht))) return done(rc);
    mark(EV_ANALYSIS); mark(EV_SYMBOLIC);
    const size_t elems = (size_t)n * (size_t)n;
    if ((rc = dalloc(&d_c, elems)) || (rc = dalloc(&d_cnt, 4))) return done(rc);
    const bool upper = mode != SPGEMM_B200_TRIPLE_REF_FULL;
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, 32, g.stream);
    if (e == cudaSuccess) e = launch_triple(lctx(), view(h), view(q), view(ht), ht->d_desc, upper, 0, n, d_c, d_cnt, env_mode("SPGEMM_B200_TRIPLE_MODE"));
    if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "triple kernel", e));
    mark(EV_NUMERIC);
    if (mode == SPGEMM_B200_TRIPLE_REF_FULL) e = launch_symmetrize(lctx(), d_c, n);
    else if (mode == SPGEMM_B200_TRIPLE_MIRROR) e = launch_mirror(lctx(), d_c, n);
    if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "triple post kernel", e));
    mark(EV_POST);
    unsigned long long* hc = reinterpret_cast<unsigned long long*>(static_cast<char*>(g.h_small) + 512);
    e = cudaMemcpyAsync(hc, d_cnt, 16, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess && elems) {
        if (mode == SPGEMM_B200_TRIPLE_UPPER) e = d2h_upper(d_c, n, c_host);
        else { e = cudaMemcpyAsync(c_host, d_c, elems * 8, cudaMemcpyDeviceToHost, g.stream); g.stats.bytes_d2h += (int64_t)elems * 8; }
    }
    mark(EV_D2H);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "triple result copy", e));
    g.stats.products = (int64_t)(hc[0] + hc[1]);
    g.stats.nnz_c = (int64_t)elems;
    g.stats.bytes_min = 2 * csr_bytes(n, h->nnz) + csr_bytes(k, q->nnz) + 8 * (int64_t)elems * (mode == 0 ? 1 : 2);
    return done(SPGEMM_B200_OK);
}

int spgemm_b200_mirror_dev(double* d_c, int n) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!d_c || n < 0) return fail(SPGEMM_B200_ERR_ARG, "mirror_dev: bad argument");
    CU(launch_mirror(lctx(), d_c, n));
    return SPGEMM_B200_OK;
}

int spgemm_b200_symmetrize_dev(double* d_c, int n) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!d_c || n < 0) return fail(SPGEMM_B200_ERR_ARG, "symmetrize_dev: bad argument");
    CU(launch_symmetrize(lctx(), d_c, n));
    return SPGEMM_B200_OK;
}

// ---- raw device buffers -----------------------------------------------------------------------------------
void* spgemm_b200_device_alloc(size_t bytes) {
    if (ensure_init()) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes ? bytes : 1, g.stream);
    if (e != cudaSuccess) { fail(SPGEMM_B200_ERR_CUDA, "device_alloc", e); return nullptr; }
    return p;
}
void spgemm_b200_device_free(void* d_ptr) {
    if (g.ready) dfree(d_ptr);
}
int spgemm_b200_copy_to_host(void* host_dst, const void* d_src, size_t bytes) {
    int rc = ensure_init();
    if (rc) return rc;
    if (bytes && (!host_dst || !d_src)) return fail(SPGEMM_B200_ERR_ARG, "copy_to_host: null pointer");
    if (bytes) CU(cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return SPGEMM_B200_OK;
}
int spgemm_b200_copy_to_device(void* d_dst, const void* host_src, size_t bytes) {
    int rc = ensure_init();
    if (rc) return rc;
    if (bytes && (!d_dst || !host_src)) return fail(SPGEMM_B200_ERR_ARG, "copy_to_device: null pointer");
    if (bytes) CU(cudaMemcpyAsync(d_dst, host_src, bytes, cudaMemcpyHostToDevice, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return SPGEMM_B200_OK;
}

int spgemm_b200_copy_upper_to_host(double* host_dst, const double* d_src, int n) {
    int rc = ensure_init();
    if (rc) return rc;
    if (n > 0 && (!host_dst || !d_src)) return fail(SPGEMM_B200_ERR_ARG, "copy_upper_to_host: null pointer");
    CU(d2h_upper(d_src, n, host_dst));
    CU(cudaStreamSynchronize(g.stream));
    return SPGEMM_B200_OK;
}

int spgemm_b200_copy_on_device(void* d_dst, const void* d_src, size_t bytes) {
    int rc = ensure_init();
    if (rc) return rc;
    if (bytes && (!d_dst || !d_src)) return fail(SPGEMM_B200_ERR_ARG, "copy_on_device: null pointer");
    if (bytes) CU(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, g.stream));
    return SPGEMM_B200_OK;
}

// ---- peer memory ---------------------------------------------------------------------------------------------
void* spgemm_b200_shared_alloc(size_t bytes) {
    if (ensure_init()) return nullptr;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(SPGEMM_B200_ERR_CUDA, "shared_alloc", e); return nullptr; }
    return p;
}
void spgemm_b200_shared_free(void* d_ptr) {
    if (g.ready && d_ptr) cudaFree(d_ptr);
}
int spgemm_b200_ipc_export(const void* d_ptr, unsigned char* handle) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!d_ptr || !handle) return fail(SPGEMM_B200_ERR_ARG, "ipc_export: null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == SPGEMM_B200_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle, &h, sizeof h);
    return SPGEMM_B200_OK;
}
int spgemm_b200_ipc_open(const unsigned char* handle, void** d_ptr) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!handle || !d_ptr) return fail(SPGEMM_B200_ERR_ARG, "ipc_open: null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SPGEMM_B200_OK;
}
int spgemm_b200_ipc_close(void* d_ptr) {
    if (!g.ready || !d_ptr) return SPGEMM_B200_OK;
    CU(cudaIpcCloseMemHandle(d_ptr));
    return SPGEMM_B200_OK;
}

// ---- stopwatch / L2 flush ------------------------------------------------------------------------------------
static cudaEvent_t t_ev0 = nullptr, t_ev1 = nullptr;
static void* g_flush_buf = nullptr;
static const size_t kFlushBytes = (size_t)512 << 20;

int spgemm_b200_timer_start(void) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!t_ev0) { CU(cudaEventCreate(&t_ev0)); CU(cudaEventCreate(&t_ev1)); }
    CU(cudaEventRecord(t_ev0, g.stream));
    return SPGEMM_B200_OK;
}
int spgemm_b200_timer_stop(double* ms) {
    if (!g.ready || !t_ev0) return fail(SPGEMM_B200_ERR_STATE, "timer_stop without timer_start");
    CU(cudaEventRecord(t_ev1, g.stream));
    CU(cudaEventSynchronize(t_ev1));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, t_ev0, t_ev1));
    if (ms) *ms = f;
    return SPGEMM_B200_OK;
}
// The write evicts everything else from L2; the read pass that follows replaces the (dirty) lines of the flush
// buffer itself by clean ones, so that the next kernel is not charged the write-back of 126 MB it never wrote.
__global__ void k_flush_read(const int4* __restrict__ p, size_t n, int* __restrict__ sink) {
    int acc = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int4 v = __ldg(p + i);
        acc ^= v.x ^ v.y ^ v.z ^ v.w;
    }
    if (acc == 0x12345678) *sink = acc;          // never true for the 0x5a pattern; keeps the loads alive
}
int spgemm_b200_flush_l2(void) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!g_flush_buf) CU(cudaMalloc(&g_flush_buf, kFlushBytes + 256));
    CU(cudaMemsetAsync(g_flush_buf, 0x5a, kFlushBytes, g.stream));
    const char* base = static_cast<const char*>(g_flush_buf);
    k_flush_read<<<g.sm_count * 8, 256, 0, g.stream>>>(reinterpret_cast<const int4*>(base + kFlushBytes / 2),
                                                       kFlushBytes / 2 / sizeof(int4),
                                                       reinterpret_cast<int*>(const_cast<char*>(base) + kFlushBytes));
    CU(cudaGetLastError());
    return SPGEMM_B200_OK;
}

// ---- row costs / partition ------------------------------------------------------------------------------
int spgemm_b200_row_costs(const spgemm_b200_mat* a, const spgemm_b200_mat* b, const spgemm_b200_mat* q, int upper_only,
                          int64_t* d_costs, int64_t* total_host) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!a || !b) return fail(SPGEMM_B200_ERR_ARG, "row_costs: null matrix");
    const int m = a->rows;
    int64_t* costs = d_costs;
    if (!costs && (rc = dalloc(&costs, (size_t)m))) return rc;
    LaunchCtx lc = lctx();
    cudaError_t e = cudaSuccess;
    int32_t* ws = nullptr;
    if (q) {
        e = launch_triple_costs(lc, view(a), view(q), view(b), upper_only != 0, costs);
    } else {
        if ((rc = dalloc(&ws, (size_t)m + (size_t)SYM_BINS * m + 32))) { if (!d_costs) dfree(costs); return rc; }
        int32_t* small = ws + m + (size_t)SYM_BINS * m;
        if (reinterpret_cast<uintptr_t>(small) & 7) small += 1;
        e = cudaMemsetAsync(small, 0, 28 * 4, g.stream);
        if (e == cudaSuccess)
            e = launch_row_products(lc, view(a), view(b), 0, m, upper_only != 0, costs, ws, ws + m, small,
                                    reinterpret_cast<unsigned long long*>(small + 24));
    }
    if (e == cudaSuccess && total_host) {
        // total on the host: copy the costs back (setup path, not timed)
        std::vector<int64_t> hc((size_t)m);
        e = cudaMemcpyAsync(hc.data(), costs, (size_t)m * 8, cudaMemcpyDeviceToHost, g.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
        int64_t t = 0;
        for (int64_t v : hc) t += v;
        *total_host = t;
    }
    dfree(ws);
    if (!d_costs) dfree(costs);
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "row_costs", e);
    return SPGEMM_B200_OK;
}

int spgemm_b200_partition(const int64_t* d_costs, int rows, int parts, int32_t* bounds_host) {
    int rc = ensure_init();
    if (rc) return rc;
    if (!d_costs || !bounds_host || rows < 0 || parts <= 0) return fail(SPGEMM_B200_ERR_ARG, "partition: bad argument");
    std::vector<int64_t> c((size_t)rows);
    if (rows) {
        CU(cudaMemcpyAsync(c.data(), d_costs, (size_t)rows * 8, cudaMemcpyDeviceToHost, g.stream));
        CU(cudaStreamSynchronize(g.stream));
    }
    // +1 per row so empty rows still spread (they cost a write of zeros / an indptr entry)
    long double total = 0;
    for (int i = 0; i < rows; ++i) total += (long double)c[i] + 1;
    bounds_host[0] = 0;
    long double acc = 0;
    int p = 1;
    for (int i = 0; i < rows && p < parts; ++i) {
        acc += (long double)c[i] + 1;
        while (p < parts && acc >= total * p / parts) bounds_host[p++] = i + 1;
    }
    while (p < parts) bounds_host[p++] = rows;
    bounds_host[parts] = rows;
    return SPGEMM_B200_OK;
}

}  // extern "C"
