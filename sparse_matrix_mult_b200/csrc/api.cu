// api.cu -- the single-device C ABI of libspgemm_b200.so (include/spgemm_b200.h): contexts, memory, orchestration.
//
// Threading: the library keeps per-call state (stats, event mask, pinned staging) in one context per device.
// Every public entry point takes that context's lock for its whole duration and makes the context's device
// current (restoring the caller's device on return), so concurrent callers -- ctypes.CDLL drops the GIL -- are
// serialised per device and never see each other's state.  The reference C library is stateless
// (/root/reference/src/sparse_sparse_sparse.cpp:172-299 allocates everything per call).
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

#include "ctx.h"

using namespace sb;
using namespace sbh;

// ===================================================================================================
// contexts
namespace sbh {

static Ctx g_ctx[kMaxDevices];
static std::mutex g_mu;                     // guards creation/destruction of contexts and g_default
static int g_default = -1;                  // device bound by spgemm_b200_init (or the first call)
static thread_local Ctx* t_ctx = nullptr;
static thread_local std::string t_err;

int fail(int code, const char* what, cudaError_t e) {
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

CallGuard::CallGuard(Ctx* c) : ctx_(c), prev_ctx_(t_ctx), prev_dev_(-1) {
    ctx_->mu.lock();
    int cur = -1;
    if (cudaGetDevice(&cur) == cudaSuccess && cur != ctx_->device) {
        prev_dev_ = cur;
        cudaSetDevice(ctx_->device);
    }
    t_ctx = ctx_;
}
CallGuard::~CallGuard() {
    t_ctx = prev_ctx_;
    if (prev_dev_ >= 0) cudaSetDevice(prev_dev_);
    ctx_->mu.unlock();
}

Ctx& cx() { return *t_ctx; }

static int ctx_create(Ctx& g, int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(SPGEMM_B200_ERR_CUDA, "no CUDA device available (libspgemm_b200 has no CPU fallback)", e);
    }
    if (device < 0 || device >= n || device >= kMaxDevices) return fail(SPGEMM_B200_ERR_ARG, "device ordinal out of range");
    int prev = -1;
    cudaGetDevice(&prev);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev == device ? -1 : prev};
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
    CU(cudaEventCreate(&g.t_ev0));
    CU(cudaEventCreate(&g.t_ev1));
    CU(cudaHostAlloc(&g.h_small, 4096, cudaHostAllocPortable));
    // A PRIVATE stream-ordered pool: freed workspaces are cached for the next call (repeated calls reuse them)
    // without touching the device's default pool, which other libraries in the process (torch) share.  The cache
    // is bounded: beyond SPGEMM_B200_POOL_KEEP_GB (default 32) freed memory goes back to the driver at the next
    // synchronisation, and spgemm_b200_trim() releases all of it.
    cudaMemPoolProps props = {};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    CU(cudaMemPoolCreate(&g.pool, &props));
    const char* keep_env = getenv("SPGEMM_B200_POOL_KEEP_GB");
    uint64_t keep = (uint64_t)(keep_env ? atoll(keep_env) : 32) << 30;
    CU(cudaMemPoolSetAttribute(g.pool, cudaMemPoolAttrReleaseThreshold, &keep));
    CU(sparse_kernels_configure());
    CU(dense_kernels_configure());
    CU(triple_kernels_configure());
    g.ready = true;
    return SPGEMM_B200_OK;
}

static void ctx_destroy(Ctx& g) {
    if (!g.ready) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.own_stream);
    for (int i = 0; i < EV_COUNT; ++i) cudaEventDestroy(g.ev[i]);
    for (int i = 0; i < 3; ++i) { cudaStreamDestroy(g.aux[i]); cudaEventDestroy(g.join_ev[i]); g.aux[i] = nullptr; }
    cudaEventDestroy(g.fork_ev);
    cudaEventDestroy(g.t_ev0);
    cudaEventDestroy(g.t_ev1);
    cudaFreeHost(g.h_small);
    if (g.flush_buf) cudaFree(g.flush_buf);
    g.flush_buf = nullptr;
    cudaStreamDestroy(g.own_stream);
    if (g.pool) cudaMemPoolDestroy(g.pool);
    g.pool = nullptr;
    g.own_stream = g.stream = nullptr;
    g.ready = false;
    if (prev >= 0 && prev != g.device) cudaSetDevice(prev);
}

Ctx* device_ctx(int device) {
    if (device < 0 || device >= kMaxDevices) { fail(SPGEMM_B200_ERR_ARG, "device ordinal out of range"); return nullptr; }
    Ctx& g = g_ctx[device];
    if (g.ready) return &g;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g.ready) return &g;
    return ctx_create(g, device) == SPGEMM_B200_OK ? &g : nullptr;
}

Ctx* live_ctx(int device) {
    if (device < 0 || device >= kMaxDevices) return nullptr;
    return g_ctx[device].ready ? &g_ctx[device] : nullptr;
}

Ctx* default_ctx() {
    int d;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_default < 0) {
            const char* env = getenv("SPGEMM_B200_DEVICE");
            g_default = env ? atoi(env) : 0;
        }
        d = g_default;
    }
    Ctx* c = device_ctx(d);
    if (!c) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_default = -1;
    }
    return c;
}

// every entry point starts with one of these
#define ENTER_DEFAULT()                                   \
    Ctx* ctx__ = default_ctx();                           \
    if (!ctx__) return SPGEMM_B200_ERR_CUDA;              \
    CallGuard guard__(ctx__)
#define ENTER_DEVICE(dev)                                 \
    Ctx* ctx__ = device_ctx(dev);                         \
    if (!ctx__) return SPGEMM_B200_ERR_CUDA;              \
    CallGuard guard__(ctx__)

LaunchCtx lctx() {
    Ctx& g = cx();
    return LaunchCtx{g.stream, g.sm_count, &g.launches, {g.aux[0], g.aux[1], g.aux[2]}, g.fork_ev,
                     {g.join_ev[0], g.join_ev[1], g.join_ev[2]}};
}

void dfree(void* p) {
    if (p) cudaFreeAsync(p, cx().stream);
}

// kernel-variant overrides for experiments: SPGEMM_B200_DENSE_MODE / SPGEMM_B200_TRIPLE_MODE
static int env_mode(const char* name) {
    const char* v = getenv(name);
    return v ? atoi(v) : 0;
}

// estimate of products per output element from the operand sizes alone (no pass over the data)
static double products_per_out(const spgemm_b200_mat* a, const spgemm_b200_mat* b) {
    const double out = (double)a->rows * (double)b->cols;
    if (out <= 0 || b->rows <= 0) return 0.0;
    return (double)a->nnz * ((double)b->nnz / (double)b->rows) / out;
}

void mark(int ev) {
    Ctx& g = cx();
    cudaEventRecord(g.ev[ev], g.stream);
    g.ev_mask |= 1u << ev;
}
// `lean` calls (device-resident entry points, which may sit inside a tight timed loop) record only the events
// around their kernels; an unrecorded phase boundary coincides with the previous recorded one.
void begin_call(bool lean) {
    Ctx& g = cx();
    g.launches = 0;
    memset(&g.stats, 0, sizeof g.stats);
    g.stats.device = g.device;
    g.ev_mask = 0;
    if (!lean) mark(EV_START);
    g.ev_pending = true;
}

// fold event times into stats (blocks until the last recorded event)
void finish_stats() {
    Ctx& g = cx();
    if (!g.ev_pending) return;
    g.ev_pending = false;
    g.stats.launches = g.launches;
    if (!g.ev_mask) return;
    int last = 0, first = EV_COUNT;
    for (int i = 0; i < EV_COUNT; ++i) if (g.ev_mask & (1u << i)) { last = i; if (first == EV_COUNT) first = i; }
    cudaEventSynchronize(g.ev[last]);
    auto phase = [&](int b) {          // time of the phase ending at event b = since the closest recorded event before it
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

// ===================================================================================================
// pinned host cache (process-wide; page-locked for every device: cudaHostAllocPortable)
namespace {
struct HostCache {
    std::mutex mu;
    std::multimap<size_t, void*> free_list;
    std::unordered_map<void*, size_t> sizes;
    std::unordered_map<void*, uint64_t> freed_at;            // cached buffers: when they came back (eviction order)
    uint64_t clock = 0;
    size_t cached = 0;
    // Keeps the buffers of the most recent results for reuse: up to `limit` bytes in total, or -- when a single
    // freed buffer is larger than that -- that one buffer alone (so a loop over the same product never pays
    // cudaHostAlloc twice, and nothing older stays locked beside it).
    size_t limit = (size_t)4 << 30;
    HostCache() {
        if (const char* v = getenv("SPGEMM_B200_PINNED_CACHE_GB")) limit = (size_t)atoll(v) << 30;
    }
} g_host;
}  // namespace

void* host_cache_alloc(size_t bytes) {
    const size_t granule = (size_t)1 << 20;
    const size_t size = ((bytes ? bytes : 1) + granule - 1) / granule * granule;
    {
        std::lock_guard<std::mutex> hl(g_host.mu);
        auto it = g_host.free_list.lower_bound(size);      // smallest cached buffer that fits and wastes < half
        if (it != g_host.free_list.end() && it->first <= 2 * size) {
            void* p = it->second;
            g_host.cached -= it->first;
            g_host.free_list.erase(it);
            g_host.freed_at.erase(p);
            return p;
        }
    }
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, size, cudaHostAllocPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        host_cache_clear();                                  // drop the cache and retry once
        e = cudaHostAlloc(&p, size, cudaHostAllocPortable);
        if (e != cudaSuccess) { fail(SPGEMM_B200_ERR_CUDA, "cudaHostAlloc", e); return nullptr; }
    }
    std::lock_guard<std::mutex> hl(g_host.mu);
    g_host.sizes[p] = size;
    return p;
}

void host_cache_free(void* p) {
    if (!p) return;
    bool any_ctx = false;
    for (int d = 0; d < kMaxDevices; ++d) any_ctx = any_ctx || g_ctx[d].ready;
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> hl(g_host.mu);
        auto it = g_host.sizes.find(p);
        if (it == g_host.sizes.end()) return;
        const size_t size = it->second;
        if (!any_ctx) {                                      // freed after spgemm_b200_shutdown: nothing to cache for
            g_host.sizes.erase(it);
            drop.push_back(p);
        } else {
            const size_t budget = size > g_host.limit ? size : g_host.limit;
            // make room: evict the buffers that came back longest ago first (the results of an earlier, different
            // product; evicting by size instead kept a large stale buffer and dropped the ones the loop reuses)
            while (!g_host.free_list.empty() && g_host.cached + size > budget) {
                auto victim = g_host.free_list.begin();
                for (auto c = g_host.free_list.begin(); c != g_host.free_list.end(); ++c)
                    if (g_host.freed_at[c->second] < g_host.freed_at[victim->second]) victim = c;
                g_host.cached -= victim->first;
                g_host.sizes.erase(victim->second);
                g_host.freed_at.erase(victim->second);
                drop.push_back(victim->second);
                g_host.free_list.erase(victim);
            }
            g_host.free_list.emplace(size, p);
            g_host.freed_at[p] = ++g_host.clock;
            g_host.cached += size;
        }
    }
    for (void* q : drop) cudaFreeHost(q);
}

void host_cache_clear() {
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> hl(g_host.mu);
        for (auto& kv : g_host.free_list) { drop.push_back(kv.second); g_host.sizes.erase(kv.second); }
        g_host.free_list.clear();
        g_host.freed_at.clear();
        g_host.cached = 0;
    }
    for (void* q : drop) cudaFreeHost(q);
}

// ===================================================================================================
// operands
int alloc_mat(int rows, int cols, int64_t nnz, spgemm_b200_mat** out) {
    Ctx& g = cx();
    spgemm_b200_mat* m = new spgemm_b200_mat{rows, cols, nnz, nullptr, nullptr, nullptr, true, g.device,
                                             nullptr, false, false, false, false, false, nullptr, false, nullptr};
    int rc;
    if ((rc = dalloc(&m->ptr, (size_t)rows + 1)) || (rc = dalloc(&m->idx, (size_t)nnz)) ||
        (rc = dalloc(&m->val, (size_t)nnz))) {
        mat_release(m);
        return rc;
    }
    *out = m;
    return SPGEMM_B200_OK;
}

int upload(int rows, int cols, const int32_t* ptr, const int32_t* idx, const double* val, spgemm_b200_mat** out) {
    Ctx& g = cx();
    const int64_t nnz = rows > 0 ? (int64_t)ptr[rows] - ptr[0] : 0;
    if (nnz < 0) return fail(SPGEMM_B200_ERR_ARG, "indptr is not non-decreasing");
    if (nnz > 0 && (!idx || !val)) return fail(SPGEMM_B200_ERR_ARG, "null indices/values with nnz > 0");
    if (rows > 0 && ptr[0] != 0) return fail(SPGEMM_B200_ERR_ARG, "indptr[0] must be 0");
    spgemm_b200_mat* m = nullptr;
    int rc = alloc_mat(rows, cols, nnz, &m);
    if (rc) return rc;
    cudaError_t e = cudaSuccess;
    if (rows > 0) e = cudaMemcpyAsync(m->ptr, ptr, ((size_t)rows + 1) * 4, cudaMemcpyHostToDevice, g.stream);
    else e = cudaMemsetAsync(m->ptr, 0, 4, g.stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->idx, idx, (size_t)nnz * 4, cudaMemcpyHostToDevice, g.stream);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->val, val, (size_t)nnz * 8, cudaMemcpyHostToDevice, g.stream);
    if (e != cudaSuccess) {
        mat_release(m);
        return fail(SPGEMM_B200_ERR_CUDA, "operand upload", e);
    }
    g.stats.bytes_h2d += csr_bytes(rows, nnz);
    *out = m;
    return SPGEMM_B200_OK;
}

static void panel_cache_drop(spgemm_b200_mat* m);

void mat_release(spgemm_b200_mat* m) {
    if (!m) return;
    panel_cache_drop(m);
    if (m->owns) { dfree(m->ptr); dfree(m->idx); dfree(m->val); }
    dfree(m->d_flags);
    if (m->shadow) mat_release(m->shadow);
    delete m;
}

void result_release(spgemm_b200_result* r) {
    if (!r) return;
    dfree(r->d_ptr); dfree(r->d_idx); dfree(r->d_val);
    delete r;
}

// Build X^T on the device; sort_desc: every row sorted by DESCENDING column (see launch_sort_rows), else rows come
// out in the arrival order of the scatter (all the window kernel of the triple product needs).
int transpose_impl(const spgemm_b200_mat* x, spgemm_b200_mat** out, bool sort_desc) {
    Ctx& g = cx();
    NvtxRange nv("spgemm_b200:transpose");
    spgemm_b200_mat* t = new spgemm_b200_mat{x->cols, x->rows, x->nnz, nullptr, nullptr, nullptr, true, g.device,
                                             nullptr, false, false, false, false, false, nullptr, false, nullptr};
    int32_t *counts = nullptr, *cursor = nullptr;
    int64_t* tmp = nullptr;
    int rc;
    if ((rc = dalloc(&t->ptr, (size_t)t->rows + 1)) || (rc = dalloc(&t->idx, (size_t)t->nnz)) ||
        (rc = dalloc(&t->val, (size_t)t->nnz)) || (rc = dalloc(&counts, (size_t)t->rows + 1)) ||
        (rc = dalloc(&cursor, (size_t)t->rows + 1)) || (rc = dalloc(&tmp, 1032))) {
        mat_release(t); dfree(counts); dfree(cursor); dfree(tmp);
        return rc;
    }
    LaunchCtx lc = lctx();
    cudaError_t e = cudaMemsetAsync(counts, 0, ((size_t)t->rows + 1) * 4, g.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(cursor, 0, ((size_t)t->rows + 1) * 4, g.stream);
    if (e == cudaSuccess) e = launch_transpose_count(lc, view(x), x->nnz, counts);
    if (e == cudaSuccess) e = launch_scan_i32(lc, counts, t->ptr, t->rows, tmp);
    if (e == cudaSuccess) e = launch_transpose_fill(lc, view(x), x->nnz, t->ptr, cursor, t->idx, t->val);
    // rows of the transpose come out in atomic order: sort them by descending column (counts is reused as the
    // list of rows too long for the per-thread sort)
    if (e == cudaSuccess && sort_desc) e = launch_sort_rows(lc, t->rows, t->ptr, t->idx, t->val, counts, true);
    t->desc_sorted = sort_desc;
    dfree(counts); dfree(cursor); dfree(tmp);
    if (e != cudaSuccess) {
        mat_release(t);
        return fail(SPGEMM_B200_ERR_CUDA, "transpose", e);
    }
    *out = t;
    return SPGEMM_B200_OK;
}

int ensure_checked(spgemm_b200_mat* m1, spgemm_b200_mat* m2) {
    Ctx& g = cx();
    spgemm_b200_mat* ms[2] = {m1, m2 == m1 ? nullptr : m2};
    int32_t* h = static_cast<int32_t*>(g.h_small) + 64;      // 2 x int32[8] of the pinned staging block
    bool pending = false;
    for (int k = 0; k < 2; ++k) {
        spgemm_b200_mat* m = ms[k];
        if (!m || m->checked) continue;
        if (!m->d_flags) {
            int rc = dalloc(&m->d_flags, 8);
            if (rc) return rc;
        }
        CU(launch_check_csr(lctx(), view(m), m->nnz, m->d_flags));
        CU(cudaMemcpyAsync(h + 8 * k, m->d_flags, 32, cudaMemcpyDeviceToHost, g.stream));
        pending = true;
    }
    if (!pending) {
        for (int k = 0; k < 2; ++k)
            if (ms[k] && !ms[k]->valid) return fail(SPGEMM_B200_ERR_ARG, "invalid CSR operand (rejected earlier)");
        return SPGEMM_B200_OK;
    }
    CU(cudaStreamSynchronize(g.stream));
    for (int k = 0; k < 2; ++k) {
        spgemm_b200_mat* m = ms[k];
        if (!m || m->checked) continue;
        m->checked = true;
        m->sorted = h[8 * k] != 0;
        m->valid = h[8 * k + 3] == 0;
        m->runs = m->sorted && h[8 * k + 4] == h[8 * k + 5];
    }
    for (int k = 0; k < 2; ++k)
        if (ms[k] && !ms[k]->valid)
            return fail(SPGEMM_B200_ERR_ARG, k == 0 ? "first operand: column index out of range or indptr not monotone"
                                                    : "second operand: column index out of range or indptr not monotone");
    return SPGEMM_B200_OK;
}

int sorted_view(spgemm_b200_mat* m, spgemm_b200_mat** out) {
    Ctx& g = cx();
    *out = m;
    if (!m->checked || m->sorted || m->nnz == 0) return SPGEMM_B200_OK;
    NvtxRange nv("spgemm_b200:canonicalise");
    spgemm_b200_mat* t = m;
    if (!m->owns) {                                           // borrowed arrays are never modified: sort a copy
        if (m->shadow) { *out = m->shadow; return SPGEMM_B200_OK; }
        t = new spgemm_b200_mat{m->rows, m->cols, m->nnz, nullptr, nullptr, nullptr, true, g.device,
                                nullptr, false, false, false, false, false, nullptr, false, nullptr};
        int rc;
        if ((rc = dalloc(&t->ptr, (size_t)m->rows + 1)) || (rc = dalloc(&t->idx, (size_t)m->nnz)) ||
            (rc = dalloc(&t->val, (size_t)m->nnz)) || (rc = dalloc(&t->d_flags, 8))) {
            mat_release(t);
            return rc;
        }
        CU(cudaMemcpyAsync(t->ptr, m->ptr, ((size_t)m->rows + 1) * 4, cudaMemcpyDeviceToDevice, g.stream));
        CU(cudaMemcpyAsync(t->idx, m->idx, (size_t)m->nnz * 4, cudaMemcpyDeviceToDevice, g.stream));
        CU(cudaMemcpyAsync(t->val, m->val, (size_t)m->nnz * 8, cudaMemcpyDeviceToDevice, g.stream));
        m->shadow = t;
    }
    int32_t* list = nullptr;
    int rc = dalloc(&list, (size_t)t->rows + 1);
    if (rc) return rc;
    cudaError_t e = launch_sort_rows(lctx(), t->rows, t->ptr, t->idx, t->val, list, false);
    const int32_t one = 1;                                    // the kernels read the flag from the device
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->d_flags, &one, 4, cudaMemcpyHostToDevice, g.stream);
    dfree(list);
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "row sort", e);
    t->checked = true;
    t->sorted = true;
    t->valid = true;
    t->desc_sorted = false;
    *out = t;
    return SPGEMM_B200_OK;
}

// ===================================================================================================
// host copies of symmetric results
// Zero `count` doubles with non-temporal stores: the destination is not read first and the zeros do not displace
// the cache.
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

// Device -> host copy of rows [r0, r1) of an n-column result whose entries left of the diagonal are known to be
// zero (symmetric dense mode, triple product): only the upper trapezoids cross PCIe, host threads zero the rest.
// The n rows are cut into fixed row blocks of kUpperBlocks-th of the matrix; block g (rows [gG, (g+1)G)) sends
// columns [gG, n) as one 2-D copy and has columns [0, gG) zeroed on the host.  Halves the bytes on the link.
// The zero fill (class ZeroFill) is work any thread can do whoever owns the rows: share `part` of `nparts` (the
// multi-GPU driver passes its GPU's index so the lower triangle is split evenly over all workers; one GPU: 0 of 1)
// takes the blocks g = part (mod nparts), on `SPGEMM_B200_ZERO_THREADS` threads (default: the host's cores / nparts,
// at most 4).  d_c holds rows [r0, r1) only; c_host is the full n-column host matrix.
constexpr int kUpperBlocks = 256;

// The zero fill runs on its own host threads from the START of a call (it needs nothing from the GPU), beside the
// operand upload, the kernels and the copy of the result.  It does not make the call shorter by much: the 16-core
// hosts of this pool sustain ~80-120 GB/s of memory traffic in total, which the zero fill (6.4 GB for cfg 5), the
// upload (0.9 GB) and the DMA writes of the result (6.4 GB) share -- measured end to end on one GPU, cfg 5:
// 170.6 / 167.4 / 177.0 ms with 2 / 4 / 8 zeroing threads (more threads slow the upload down: 16 -> 38 ms), against
// 169-173 ms when the zero fill only started with the copy (profiles/r2/e2e_probe_early_zero_fill.jsonl).
void ZeroFill::start(double* c_host, int n, int part, int nparts) {
    if (n <= 0 || !c_host) return;
    // (the copy alone takes 123 ms on one GPU -- PCIe at 52 GB/s.  On 2 - 8 GPUs copy + zero fill take ~105 ms whatever
    //  the thread count: 12.8 GB written into host memory at ~120 GB/s, the host's memory write bandwidth)
    unsigned hw = std::thread::hardware_concurrency();
    int nthreads = hw == 0 ? 4 : (int)hw / (nparts > 0 ? nparts : 1);
    if (nthreads > 4) nthreads = 4;
    if (nthreads < 1) nthreads = 1;
    if (const char* ev = getenv("SPGEMM_B200_ZERO_THREADS")) nthreads = atoi(ev);      // 0 = skip (experiments only)
    if ((size_t)n * n < ((size_t)1 << 22) && nthreads > 1) nthreads = 1;
    if (nthreads <= 0) return;
    const int G = (n + kUpperBlocks - 1) / kUpperBlocks;
    const int nblocks = (n + G - 1) / G;
    auto zero_blocks = [=](int t) {
        // blocks of this share, dealt to the threads from the bottom of the matrix up (widest rectangles first)
        int k = 0;
        for (int gb = nblocks - 1; gb >= 1; --gb) {
            if (gb % nparts != part) continue;
            if (k++ % nthreads != t) continue;
            const int b0 = gb * G, b1 = b0 + G < n ? b0 + G : n;
            for (int r = b0; r < b1; ++r) zero_nt(c_host + (size_t)r * n, (size_t)b0);
        }
        _mm_sfence();
    };
    for (int t = 0; t < nthreads; ++t) threads_.emplace_back(zero_blocks, t);
}
void ZeroFill::join() {
    for (auto& th : threads_) th.join();
    threads_.clear();
}

cudaError_t d2h_upper_rows(const double* d_c, int n, int r0, int r1, double* c_host) {
    Ctx& g = cx();
    if (n <= 0) return cudaSuccess;
    const int G = (n + kUpperBlocks - 1) / kUpperBlocks;          // rows per block
    cudaError_t err = cudaSuccess;
    for (int b0 = (r0 / G) * G; b0 < r1 && err == cudaSuccess; b0 += G) {
        const int s0 = b0 > r0 ? b0 : r0, s1 = b0 + G < r1 ? b0 + G : r1;     // this call's rows of the block
        if (s1 <= s0) continue;
        err = cudaMemcpy2DAsync(c_host + (size_t)s0 * n + b0, (size_t)n * 8, d_c + (size_t)(s0 - r0) * n + b0,
                                (size_t)n * 8, (size_t)(n - b0) * 8, (size_t)(s1 - s0), cudaMemcpyDeviceToHost, g.stream);
        g.stats.bytes_d2h += (int64_t)(n - b0) * (s1 - s0) * 8;
    }
    return err;
}

// ===================================================================================================
// compute building blocks

// Sparse product of rows [r0, r1).  Records EV_ANALYSIS / EV_SYMBOLIC / EV_NUMERIC.
int csr_impl(spgemm_b200_mat* a, spgemm_b200_mat* b_in, int upper_only, int r0, int r1, spgemm_b200_result** out) {
    Ctx& g = cx();
    const int m = r1 - r0, n = b_in->cols;
    int rc;
    if ((rc = ensure_checked(a, b_in))) return rc;
    spgemm_b200_mat* b = b_in;
    if ((rc = sorted_view(b_in, &b))) return rc;
    spgemm_b200_result* res = new spgemm_b200_result{m, n, 0, nullptr, nullptr, nullptr, g.device};
    rc = dalloc(&res->d_ptr, (size_t)m + 1);
    if (rc) { delete res; return rc; }
    g.stats.bytes_min = csr_bytes(m, (int64_t)0) + csr_bytes(b->rows, b->nnz);
    if (m == 0 || a->nnz == 0 || b->nnz == 0) {
        CU(cudaMemsetAsync(res->d_ptr, 0, ((size_t)m + 1) * 8, g.stream));
        mark(EV_ANALYSIS); mark(EV_SYMBOLIC); mark(EV_NUMERIC);
        rc = dalloc(&res->d_idx, 1);
        if (!rc) rc = dalloc(&res->d_val, 1);
        if (rc) { result_release(res); return rc; }
        *out = res;
        return SPGEMM_B200_OK;
    }

    // one workspace block: nnz[m] | lists[BINS*m] | small counters | scan scratch
    const int bins = SYM_BINS > NUM_BINS ? SYM_BINS : NUM_BINS;
    const size_t small_ints = 32;    // cursor[16] | work[8] | pad ; total (u64) lives at small + 24
    const size_t ws_ints = (size_t)m + (size_t)bins * m + small_ints;
    int32_t* ws = nullptr;
    int64_t* scan_tmp = nullptr;
    if ((rc = dalloc(&ws, ws_ints)) || (rc = dalloc(&scan_tmp, 1032))) {
        dfree(ws); result_release(res);
        return rc;
    }
    int32_t* d_nnz = ws;
    int32_t* d_lists = ws + m;
    int32_t* d_small = d_lists + (size_t)bins * m;
    if ((reinterpret_cast<uintptr_t>(d_small) & 7) != 0) d_small += 1;   // keep the 8-byte total aligned (slack above)
    int32_t* d_cursor = d_small;
    int32_t* d_work = d_small + 16;
    unsigned long long* d_total = reinterpret_cast<unsigned long long*>(d_small + 24);
    auto bail = [&](int code) {
        dfree(ws); dfree(scan_tmp); result_release(res);
        return code;
    };
    LaunchCtx lc = lctx();
    SparseJob job{view(a), view(b), r0, m, upper_only != 0, b->d_flags};
    cudaError_t e;
    {
        NvtxRange nv("spgemm_b200:analysis");
        e = cudaMemsetAsync(d_small, 0, 28 * sizeof(int32_t), g.stream);
        if (e == cudaSuccess)
            e = launch_row_products(lc, job.A, job.B, r0, m, job.upper_only, nullptr, d_nnz, d_lists, d_cursor, d_total);
        if (e != cudaSuccess) return bail(fail(SPGEMM_B200_ERR_CUDA, "row products", e));
        mark(EV_ANALYSIS);
    }
    int32_t* h = static_cast<int32_t*>(g.h_small);
    e = cudaMemcpyAsync(h, d_small, 28 * sizeof(int32_t), cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) return bail(fail(SPGEMM_B200_ERR_CUDA, "bin counts", e));
    int32_t sym_counts[SYM_BINS];
    for (int k = 0; k < SYM_BINS; ++k) sym_counts[k] = h[k];
    unsigned long long total_products;
    memcpy(&total_products, h + 24, 8);
    g.stats.products = (int64_t)total_products;

    // Heavy rows: the symbolic phase keeps their column bitmaps (up to SPGEMM_B200_BITMAP_KEEP_MB, default 4096) so the
    // numeric rank kernel does not rebuild them.  One slot per row of the bitmap bin while they last.
    SavedBitmaps saved{nullptr, nullptr, d_work + 4, 0, saved_bitmap_words(n)};
    int32_t* d_slot_of_row = nullptr;
    if (sym_counts[SYM_BITMAP] > 0) {
        const char* kv = getenv("SPGEMM_B200_BITMAP_KEEP_MB");
        const int64_t budget = (int64_t)(kv ? atoll(kv) : 4096) << 20;
        int64_t slots = budget / ((int64_t)saved.words * 4);
        if (slots > sym_counts[SYM_BITMAP]) slots = sym_counts[SYM_BITMAP];
        if (slots > 0 && dalloc(&d_slot_of_row, (size_t)m) == SPGEMM_B200_OK) {
            if (dalloc(&saved.bits, (size_t)slots * saved.words) == SPGEMM_B200_OK) {
                saved.slots = (int)slots;
                saved.slot_of_row = d_slot_of_row;
                e = cudaMemsetAsync(d_slot_of_row, 0xff, (size_t)m * 4, g.stream);
                if (e != cudaSuccess) return bail(fail(SPGEMM_B200_ERR_CUDA, "bitmap slots", e));
            } else {
                cudaGetLastError();                            // no room: the numeric phase rebuilds the bitmaps
                saved.bits = nullptr;
            }
        }
    }
    auto bail2 = [&](int code) {
        dfree(saved.bits); dfree(d_slot_of_row);
        return bail(code);
    };
    {
        NvtxRange nv("spgemm_b200:symbolic");
        e = launch_symbolic(lc, job, d_lists, sym_counts, d_nnz, d_work, saved);
        if (e == cudaSuccess) e = launch_scan_i64(lc, d_nnz, res->d_ptr, m, scan_tmp);
        if (e == cudaSuccess) e = cudaMemsetAsync(d_cursor, 0, 16 * sizeof(int32_t), g.stream);
        if (e == cudaSuccess) e = launch_bin_by_nnz(lc, d_nnz, m, d_lists, d_cursor);
        if (e != cudaSuccess) return bail2(fail(SPGEMM_B200_ERR_CUDA, "symbolic phase", e));
        mark(EV_SYMBOLIC);
    }
    int64_t* h64 = reinterpret_cast<int64_t*>(h + 32);
    e = cudaMemcpyAsync(h, d_cursor, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(h64, res->d_ptr + m, 8, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) return bail2(fail(SPGEMM_B200_ERR_CUDA, "nnz(C)", e));
    int32_t num_counts[NUM_BINS];
    for (int k = 0; k < NUM_BINS; ++k) num_counts[k] = h[k];
    res->nnz = *h64;
    g.stats.nnz_c = res->nnz;
    // bytes(A) + bytes(B) + bytes(C); a row slice of A is charged pro rata
    g.stats.bytes_min = csr_bytes(m, a->rows ? a->nnz * m / a->rows : 0) + csr_bytes(b->rows, b->nnz) + csr_bytes(m, res->nnz);
    if ((rc = dalloc(&res->d_idx, (size_t)res->nnz)) || (rc = dalloc(&res->d_val, (size_t)res->nnz))) return bail2(rc);
    {
        NvtxRange nv("spgemm_b200:numeric");
        e = launch_numeric(lc, job, d_lists, num_counts, res->d_ptr, res->d_idx, res->d_val, d_work, saved);
        if (e != cudaSuccess) return bail2(fail(SPGEMM_B200_ERR_CUDA, "numeric phase", e));
        mark(EV_NUMERIC);
    }
    dfree(saved.bits); dfree(d_slot_of_row);
    dfree(ws); dfree(scan_tmp);
    *out = res;
    return SPGEMM_B200_OK;
}

// rows [r0, r1) of the dense product into d_c ((r1-r0) x n).  Operands must have been checked.
int dense_rows(spgemm_b200_mat* a, spgemm_b200_mat* b_in, int upper_only, int r0, int r1, double* d_c) {
    Ctx& g = cx();
    spgemm_b200_mat* b = b_in;
    int rc = sorted_view(b_in, &b);
    if (rc) return rc;
    NvtxRange nv("spgemm_b200:dense");
    cudaError_t de = launch_dense(lctx(), view(a), view(b), b->d_flags, upper_only != 0, r0, r1 - r0, d_c,
                                  env_mode("SPGEMM_B200_DENSE_MODE"), products_per_out(a, b));
    if (de != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "dense kernel", de);
    g.stats.nnz_c = (int64_t)(r1 - r0) * b->cols;
    g.stats.bytes_min = csr_bytes(a->rows, a->nnz) + csr_bytes(b->rows, b->nnz) + 8 * g.stats.nnz_c;
    return SPGEMM_B200_OK;
}

// Paneled transpose of rows [plan.k0, n) of H (analysis.cu): one CSR with plan.np * H.cols rows.
struct PanelT {
    int32_t* ptr = nullptr;
    uint32_t* pk = nullptr;        // packed (row of H within its panel, low bits of the column) of every entry
    double* val = nullptr;
};
static void panels_release(PanelT& t) {
    dfree(t.ptr); dfree(t.pk); dfree(t.val);
    t = PanelT();
}
// A paneled transpose kept on its matrix (spgemm_b200_mat_cache_transpose): a caller that multiplies with the same H
// again and again (the iterations of an inversion) pays for it once.
struct PanelCache {
    PanelT t;
    TriplePlan plan;
};
static void panel_cache_drop(spgemm_b200_mat* m) {
    if (!m->panel_cache) return;
    PanelCache* c = static_cast<PanelCache*>(m->panel_cache);
    panels_release(c->t);
    delete c;
    m->panel_cache = nullptr;
}

static int transpose_panels(const spgemm_b200_mat* h, const TriplePlan& plan, PanelT* out) {
    Ctx& g = cx();
    NvtxRange nv("spgemm_b200:transpose_panels");
    const size_t trows = (size_t)plan.np * (size_t)h->cols;
    if (trows + 1 > 0x7fffffffULL) return fail(SPGEMM_B200_ERR_OVERFLOW, "triple: panels x columns of H exceed 2^31");
    if (plan.panel_w > kPanelMaxWidth) return fail(SPGEMM_B200_ERR_OVERFLOW, "triple: panel wider than the entry packing");
    PanelT t;
    int32_t *counts = nullptr, *cursor = nullptr;
    int64_t* tmp = nullptr;
    int rc;
    if ((rc = dalloc(&t.ptr, trows + 1)) || (rc = dalloc(&t.pk, (size_t)h->nnz)) || (rc = dalloc(&t.val, (size_t)h->nnz)) ||
        (rc = dalloc(&counts, trows + 1)) || (rc = dalloc(&cursor, trows + 1)) || (rc = dalloc(&tmp, 1032))) {
        panels_release(t); dfree(counts); dfree(cursor); dfree(tmp);
        return rc;
    }
    LaunchCtx lc = lctx();
    cudaError_t e = cudaMemsetAsync(counts, 0, (trows + 1) * 4, g.stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(cursor, 0, (trows + 1) * 4, g.stream);
    if (e == cudaSuccess) e = launch_transpose_count_panels(lc, view(h), h->nnz, plan.k0, h->rows, plan.panel_w, counts);
    if (e == cudaSuccess) e = launch_scan_i32(lc, counts, t.ptr, (int)trows, tmp);
    if (e == cudaSuccess)
        e = launch_transpose_fill_panels(lc, view(h), h->nnz, plan.k0, h->rows, plan.panel_w, t.ptr, cursor, t.pk, t.val);
    dfree(counts); dfree(cursor); dfree(tmp);
    if (e != cudaSuccess) {
        panels_release(t);
        return fail(SPGEMM_B200_ERR_CUDA, "paneled transpose", e);
    }
    *out = t;
    return SPGEMM_B200_OK;
}

// rows [r0, r1) of H Q H^T into d_c ((r1-r0) x n); d_cnt: device u64[4], zeroed here (P1, P2, ticket, spare).
int triple_rows(const spgemm_b200_mat* h, const spgemm_b200_mat* q, const spgemm_b200_mat* ht, int upper_only, int r0,
                int r1, double* d_c, unsigned long long* d_cnt) {
    Ctx& g = cx();
    (void)ht;      // a plain CSR transpose cannot stand in for the panels: they store packed (row, column) words
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, 32, g.stream);
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "triple counters", e);
    int rc;
    {
        TriplePlan plan = triple_plan(h->rows, r0, upper_only != 0, h->nnz, h->cols);
        PanelT t;
        spgemm_b200_mat* hm = const_cast<spgemm_b200_mat*>(h);       // the kept transpose is a cache on the handle
        PanelCache* kept = static_cast<PanelCache*>(hm->panel_cache);
        if (kept && (kept->plan.k0 != plan.k0 || kept->plan.np != plan.np || kept->plan.panel_w != plan.panel_w)) {
            panel_cache_drop(hm);
            kept = nullptr;
        }
        if (kept) t = kept->t;
        else if ((rc = transpose_panels(h, plan, &t))) return rc;
        if (!kept && hm->cache_panels) {
            kept = new PanelCache{t, plan};
            hm->panel_cache = kept;
        }
        mark(EV_ANALYSIS); mark(EV_SYMBOLIC);
        NvtxRange nv("spgemm_b200:triple");
        const bool q_runs = q->checked && q->runs && env_mode("SPGEMM_B200_TRIPLE_GENERIC") == 0;
        // Per-entry (start of the row of Q, its length, its first column), prepared once per call when rows take part
        // in several (panel, row) items: cfg 5 (4 panels) 13.46 -> 12.41 ms; with one panel the pass does not pay
        // (cfg 3: 0.471 -> 0.485 ms), the items gather the three values themselves.
        int4* entry_meta = nullptr;
        if (q_runs && h->nnz > 0 && plan.np >= 2 && (rc = dalloc(&entry_meta, (size_t)h->nnz))) {
            if (!kept) panels_release(t);
            return rc;
        }
        e = launch_triple_panels(lctx(), view(h), view(q), q_runs, t.ptr, t.pk, t.val, plan, upper_only != 0, r0, r1 - r0,
                                 d_c, d_cnt, entry_meta);
        dfree(entry_meta);
        if (!kept) panels_release(t);
    }
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "triple kernel", e);
    g.stats.nnz_c = (int64_t)(r1 - r0) * h->rows;
    g.stats.bytes_min = 2 * csr_bytes(h->rows, h->nnz) + csr_bytes(q->rows, q->nnz) + 8 * g.stats.nnz_c;
    return SPGEMM_B200_OK;
}

int row_costs_impl(const spgemm_b200_mat* a, const spgemm_b200_mat* b, const spgemm_b200_mat* q, int upper_only,
                   int dense_cols, int64_t* costs) {
    Ctx& g = cx();
    const int m = a->rows;
    LaunchCtx lc = lctx();
    cudaError_t e = cudaSuccess;
    int32_t* ws = nullptr;
    if (q) {
        const TriplePlan plan = triple_plan(a->rows, 0, upper_only != 0, a->nnz, a->cols);
        e = launch_triple_costs(lc, view(a), view(q), view(b), upper_only != 0, q->checked && q->runs, plan.np,
                                plan.panel_w, costs);
    } else {
        int rc = dalloc(&ws, (size_t)m + (size_t)SYM_BINS * m + 32);
        if (rc) return rc;
        int32_t* small = ws + m + (size_t)SYM_BINS * m;
        if (reinterpret_cast<uintptr_t>(small) & 7) small += 1;
        e = cudaMemsetAsync(small, 0, 28 * 4, g.stream);
        if (e == cudaSuccess)
            e = launch_row_products(lc, view(a), view(b), 0, m, upper_only != 0, costs, ws, ws + m, small,
                                    reinterpret_cast<unsigned long long*>(small + 24));
        // dense output: a row of `dense_cols` doubles is written whatever its products (one product ~ 3 doubles written:
        // 5 ps per L2 reduction against 1.6 ps per streamed double)
        if (e == cudaSuccess && dense_cols > 0) e = launch_add_const(lc, costs, m, (long long)(0.3 * dense_cols) + 1);
    }
    dfree(ws);
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "row_costs", e);
    return SPGEMM_B200_OK;
}

void partition_costs(const int64_t* c, int rows, int parts, int32_t* bounds) {
    // +1 per row so empty rows still spread (they cost a write of zeros / an indptr entry)
    long double total = 0;
    for (int i = 0; i < rows; ++i) total += (long double)c[i] + 1;
    bounds[0] = 0;
    long double acc = 0;
    int p = 1;
    for (int i = 0; i < rows && p < parts; ++i) {
        acc += (long double)c[i] + 1;
        while (p < parts && acc >= total * p / parts) bounds[p++] = i + 1;
    }
    while (p < parts) bounds[p++] = rows;
    bounds[parts] = rows;
}

// Bottleneck partition with a start-dependent fixed cost: block p = rows [b_p, b_{p+1}) costs
// tail_coeff * tail[b_p] + sum of its rows' costs, where tail[i] = units of rows i.. (a suffix sum; for the triple product
// the entries of H from row i on, which the rank that starts at row i has to transpose).  Minimises the largest block
// cost: binary search on the bound, greedy feasibility (taking as many rows as fit is optimal because the fixed cost
// does not grow with the start row).
void partition_costs_tail(const int64_t* c, const int64_t* tail, double tail_coeff, int rows, int parts, int32_t* bounds) {
    std::vector<long double> cum((size_t)rows + 1, 0);
    for (int i = 0; i < rows; ++i) cum[i + 1] = cum[i] + (long double)c[i] + 1;
    auto fixed = [&](int r0) { return r0 < rows ? (long double)tail_coeff * (long double)tail[r0] : 0.0L; };
    auto greedy = [&](long double limit, int32_t* out) {
        int r0 = 0;
        for (int p = 0; p < parts; ++p) {
            if (out) out[p] = r0;
            if (r0 >= rows) continue;
            const long double room = limit - fixed(r0);
            if (room <= 0) return false;
            int lo = r0, hi = rows;                       // largest r1 with cum[r1] - cum[r0] <= room
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (cum[mid] - cum[r0] <= room) lo = mid; else hi = mid - 1;
            }
            if (lo == r0) return false;                   // not even one row fits
            r0 = lo;
        }
        if (out) out[parts] = rows;
        return r0 >= rows;
    };
    long double lo = 0, hi = cum[rows] + fixed(0) + 1;
    for (int it = 0; it < 60; ++it) {
        const long double mid = (lo + hi) / 2;
        if (greedy(mid, nullptr)) hi = mid; else lo = mid;
    }
    if (!greedy(hi, bounds)) partition_costs(c, rows, parts, bounds);      // cannot happen; keep a valid answer
    bounds[parts] = rows;
}

}  // namespace sbh

static int check_csr_args(int rows, int cols, const int32_t* ptr, const char* name) {
    if (rows < 0 || cols < 0 || !ptr) return fail(SPGEMM_B200_ERR_ARG, name);
    return SPGEMM_B200_OK;
}

// ===================================================================================================
extern "C" {

const char* spgemm_b200_version(void) { return "spgemm_b200 0.2 (sm_100a)"; }

const char* spgemm_b200_last_error(void) { return t_err.c_str(); }

int spgemm_b200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int spgemm_b200_init(int device) {
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (g_default >= 0 && g_default != device && g_ctx[g_default].ready)
            return fail(SPGEMM_B200_ERR_STATE, "already initialised on another device; call spgemm_b200_shutdown first");
        g_default = device;
    }
    if (!device_ctx(device)) {
        std::lock_guard<std::mutex> lk(g_mu);
        g_default = -1;
        return t_err.find("out of range") != std::string::npos ? SPGEMM_B200_ERR_ARG : SPGEMM_B200_ERR_CUDA;
    }
    return SPGEMM_B200_OK;
}

void spgemm_b200_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (int d = 0; d < kMaxDevices; ++d) {
        if (!g_ctx[d].ready) continue;
        std::lock_guard<std::recursive_mutex> cl(g_ctx[d].mu);
        ctx_destroy(g_ctx[d]);
    }
    host_cache_clear();
    g_default = -1;
}

int spgemm_b200_trim(size_t keep_bytes) {
    for (int d = 0; d < kMaxDevices; ++d) {
        if (!g_ctx[d].ready) continue;
        CallGuard guard(&g_ctx[d]);
        CU(cudaStreamSynchronize(g_ctx[d].stream));
        CU(cudaMemPoolTrimTo(g_ctx[d].pool, keep_bytes));
    }
    host_cache_clear();
    return SPGEMM_B200_OK;
}

int spgemm_b200_get_stats(spgemm_b200_stats* out) {
    if (!out) return fail(SPGEMM_B200_ERR_ARG, "null stats");
    ENTER_DEFAULT();
    finish_stats();
    *out = cx().stats;
    return SPGEMM_B200_OK;
}

int spgemm_b200_set_stream(void* stream) {
    ENTER_DEFAULT();
    cx().stream = stream ? static_cast<cudaStream_t>(stream) : cx().own_stream;
    return SPGEMM_B200_OK;
}

int spgemm_b200_synchronize(void) {
    ENTER_DEFAULT();
    CU(cudaStreamSynchronize(cx().stream));
    return SPGEMM_B200_OK;
}

// ---- pinned host cache ------------------------------------------------------------------------------
void* spgemm_b200_host_alloc(size_t bytes) {
    Ctx* c = default_ctx();
    if (!c) return nullptr;
    CallGuard guard(c);
    return host_cache_alloc(bytes);
}

void spgemm_b200_host_free(void* p) { host_cache_free(p); }

// ---- matrices -----------------------------------------------------------------------------------------
int spgemm_b200_mat_upload(int rows, int cols, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                           const double* values, spgemm_b200_mat** out) {
    if (!out) return fail(SPGEMM_B200_ERR_ARG, "null out");
    int rc;
    if ((rc = check_csr_args(rows, cols, indptr, "mat_upload: bad matrix"))) return rc;
    if (rows > 0 && (int64_t)indptr[rows] != nnz) return fail(SPGEMM_B200_ERR_ARG, "mat_upload: nnz != indptr[rows]");
    ENTER_DEFAULT();
    return upload(rows, cols, indptr, indices, values, out);
}

int spgemm_b200_mat_wrap(int rows, int cols, int64_t nnz, const int32_t* d_indptr, const int32_t* d_indices,
                         const double* d_values, spgemm_b200_mat** out) {
    if (!out || !d_indptr || rows < 0 || cols < 0 || nnz < 0) return fail(SPGEMM_B200_ERR_ARG, "mat_wrap: bad argument");
    ENTER_DEFAULT();
    *out = new spgemm_b200_mat{rows, cols, nnz, const_cast<int32_t*>(d_indptr), const_cast<int32_t*>(d_indices),
                               const_cast<double*>(d_values), false, cx().device, nullptr, false, false, false, false, false, nullptr, false, nullptr};
    return SPGEMM_B200_OK;
}

int spgemm_b200_mat_transpose(const spgemm_b200_mat* x, spgemm_b200_mat** out) {
    if (!x || !out) return fail(SPGEMM_B200_ERR_ARG, "mat_transpose: null argument");
    ENTER_DEVICE(x->device);
    int rc = ensure_checked(const_cast<spgemm_b200_mat*>(x));
    if (rc) return rc;
    return transpose_impl(x, out, true);
}

int spgemm_b200_mat_cache_transpose(spgemm_b200_mat* x, int enable) {
    if (!x) return fail(SPGEMM_B200_ERR_ARG, "mat_cache_transpose: null argument");
    ENTER_DEVICE(x->device);
    x->cache_panels = enable != 0;
    if (!enable) panel_cache_drop(x);
    return SPGEMM_B200_OK;
}

int spgemm_b200_mat_sort(spgemm_b200_mat* x) {
    if (!x) return fail(SPGEMM_B200_ERR_ARG, "mat_sort: null argument");
    ENTER_DEVICE(x->device);
    int rc = ensure_checked(x);
    if (rc) return rc;
    spgemm_b200_mat* v = nullptr;
    return sorted_view(x, &v);
}

int spgemm_b200_mat_is_sorted(const spgemm_b200_mat* x) {
    if (!x) return -1;
    ENTER_DEVICE(x->device);
    if (ensure_checked(const_cast<spgemm_b200_mat*>(x))) return -1;
    return x->sorted ? 1 : 0;
}

void spgemm_b200_mat_free(spgemm_b200_mat* m) {
    if (!m) return;
    Ctx* c = live_ctx(m->device);
    if (!c) return;                 // after spgemm_b200_shutdown the pool (and with it the arrays) is gone
    CallGuard guard(c);
    mat_release(m);
}

// ---- sparse output --------------------------------------------------------------------------------------
int spgemm_b200_csr_dev(const spgemm_b200_mat* a, const spgemm_b200_mat* b, int upper_only, int row_begin, int row_end,
                        spgemm_b200_result** out) {
    if (!a || !b || !out) return fail(SPGEMM_B200_ERR_ARG, "csr_dev: null argument");
    if (a->cols != b->rows) return fail(SPGEMM_B200_ERR_ARG, "csr_dev: inner dimensions differ");
    if (a->device != b->device) return fail(SPGEMM_B200_ERR_ARG, "csr_dev: operands live on different devices");
    if (row_end < 0) { row_begin = 0; row_end = a->rows; }
    if (row_begin < 0 || row_end > a->rows || row_begin > row_end) return fail(SPGEMM_B200_ERR_ARG, "csr_dev: bad row range");
    ENTER_DEVICE(a->device);
    begin_call();
    mark(EV_H2D);
    int rc = csr_impl(const_cast<spgemm_b200_mat*>(a), const_cast<spgemm_b200_mat*>(b), upper_only, row_begin, row_end, out);
    mark(EV_POST); mark(EV_D2H);
    return rc;
}

int spgemm_b200_csr(int m, int k, int n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_values,
                    const int32_t* b_indptr, const int32_t* b_indices, const double* b_values, int upper_only,
                    spgemm_b200_result** out) {
    if (!out) return fail(SPGEMM_B200_ERR_ARG, "csr: null out");
    int rc;
    if ((rc = check_csr_args(m, k, a_indptr, "csr: bad A"))) return rc;
    if ((rc = check_csr_args(k, n, b_indptr, "csr: bad B"))) return rc;
    ENTER_DEFAULT();
    begin_call();
    spgemm_b200_mat *a = nullptr, *b = nullptr;
    {
        NvtxRange nv("spgemm_b200:h2d");
        if ((rc = upload(m, k, a_indptr, a_indices, a_values, &a))) return rc;
        const bool same = (a_indptr == b_indptr && a_indices == b_indices && a_values == b_values && m == k && k == n);
        if (same) b = a;
        else if ((rc = upload(k, n, b_indptr, b_indices, b_values, &b))) { mat_release(a); return rc; }
        mark(EV_H2D);
    }
    rc = csr_impl(a, b, upper_only, 0, m, out);
    mark(EV_POST); mark(EV_D2H);
    if (b != a) mat_release(b);
    mat_release(a);
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
    if (!r || !indptr) return fail(SPGEMM_B200_ERR_ARG, "result_copy: null argument");
    if (r->nnz > 0 && (!indices || !values)) return fail(SPGEMM_B200_ERR_ARG, "result_copy: null indices/values");
    if (!index64 && r->nnz > 0x7fffffffLL) return fail(SPGEMM_B200_ERR_OVERFLOW, "nnz(C) >= 2^31 needs index64");
    ENTER_DEVICE(r->device);
    Ctx& g = cx();
    NvtxRange nv("spgemm_b200:d2h");
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
    Ctx* c = live_ctx(r->device);
    if (!c) return;
    CallGuard guard(c);
    result_release(r);
}

// ---- dense output ---------------------------------------------------------------------------------------
int spgemm_b200_dense_dev(const spgemm_b200_mat* a, const spgemm_b200_mat* b, int upper_only, int row_begin, int row_end,
                          double* d_c) {
    if (!a || !b || !d_c) return fail(SPGEMM_B200_ERR_ARG, "dense_dev: null argument");
    if (a->cols != b->rows) return fail(SPGEMM_B200_ERR_ARG, "dense_dev: inner dimensions differ");
    if (a->device != b->device) return fail(SPGEMM_B200_ERR_ARG, "dense_dev: operands live on different devices");
    if (row_end < 0) { row_begin = 0; row_end = a->rows; }
    if (row_begin < 0 || row_end > a->rows || row_begin > row_end) return fail(SPGEMM_B200_ERR_ARG, "dense_dev: bad row range");
    ENTER_DEVICE(a->device);
    begin_call(true);
    int rc;
    if (!a->checked || !b->checked) {
        mark(EV_H2D);
        if ((rc = ensure_checked(const_cast<spgemm_b200_mat*>(a), const_cast<spgemm_b200_mat*>(b)))) return rc;
    }
    mark(EV_SYMBOLIC);
    if ((rc = dense_rows(const_cast<spgemm_b200_mat*>(a), const_cast<spgemm_b200_mat*>(b), upper_only, row_begin, row_end, d_c)))
        return rc;
    mark(EV_NUMERIC);
    return SPGEMM_B200_OK;
}

int spgemm_b200_dense(int m, int k, int n, const int32_t* a_indptr, const int32_t* a_indices, const double* a_values,
                      const int32_t* b_indptr, const int32_t* b_indices, const double* b_values, int upper_only,
                      int mirror, double* c_host) {
    if (!c_host && (int64_t)m * n > 0) return fail(SPGEMM_B200_ERR_ARG, "dense: null output");
    int rc;
    if ((rc = check_csr_args(m, k, a_indptr, "dense: bad A"))) return rc;
    if ((rc = check_csr_args(k, n, b_indptr, "dense: bad B"))) return rc;
    if (mirror && (!upper_only || m != n)) return fail(SPGEMM_B200_ERR_ARG, "dense: mirror needs upper_only and a square result");
    ENTER_DEFAULT();
    Ctx& g = cx();
    begin_call();
    spgemm_b200_mat *a = nullptr, *b = nullptr;
    double* d_c = nullptr;
    ZeroFill zero;                                 // symmetric result: the host zeroes the lower triangle meanwhile
    if (upper_only && !mirror && m == n) zero.start(c_host, n, 0, 1);
    auto done = [&](int code) {
        zero.join();
        dfree(d_c); mat_release(a); mat_release(b);
        return code;
    };
    {
        NvtxRange nv("spgemm_b200:h2d");
        if ((rc = upload(m, k, a_indptr, a_indices, a_values, &a))) return done(rc);
        if ((rc = upload(k, n, b_indptr, b_indices, b_values, &b))) return done(rc);
        mark(EV_H2D);
    }
    if ((rc = ensure_checked(a, b))) return done(rc);
    mark(EV_ANALYSIS); mark(EV_SYMBOLIC);
    const size_t elems = (size_t)m * (size_t)n;
    if ((rc = dalloc(&d_c, elems))) return done(rc);
    if ((rc = dense_rows(a, b, upper_only, 0, m, d_c))) return done(rc);
    mark(EV_NUMERIC);
    cudaError_t e = cudaSuccess;
    if (mirror) {
        e = launch_mirror(lctx(), d_c, n);
        if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "mirror kernel", e));
    }
    mark(EV_POST);
    if (elems) {
        NvtxRange nv("spgemm_b200:d2h");
        if (upper_only && !mirror && m == n) e = d2h_upper_rows(d_c, n, 0, n, c_host);
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
    if (!h || !q || !d_c) return fail(SPGEMM_B200_ERR_ARG, "triple_dev: null argument");
    if (h->cols != q->rows || q->rows != q->cols) return fail(SPGEMM_B200_ERR_ARG, "triple_dev: Q must be square with H.cols rows");
    if (ht && (ht->rows != h->cols || ht->cols != h->rows || ht->nnz != h->nnz))
        return fail(SPGEMM_B200_ERR_ARG, "triple_dev: ht is not the transpose of h");
    if (h->device != q->device || (ht && ht->device != h->device))
        return fail(SPGEMM_B200_ERR_ARG, "triple_dev: operands live on different devices");
    if (row_end < 0) { row_begin = 0; row_end = h->rows; }
    if (row_begin < 0 || row_end > h->rows || row_begin > row_end) return fail(SPGEMM_B200_ERR_ARG, "triple_dev: bad row range");
    ENTER_DEVICE(h->device);
    Ctx& g = cx();
    begin_call();
    mark(EV_H2D);
    int rc;
    if ((rc = ensure_checked(const_cast<spgemm_b200_mat*>(h), const_cast<spgemm_b200_mat*>(q)))) return rc;
    if (ht && (rc = ensure_checked(const_cast<spgemm_b200_mat*>(ht)))) return rc;
    unsigned long long* d_cnt = nullptr;
    if ((rc = dalloc(&d_cnt, 4))) return rc;
    rc = triple_rows(h, q, ht, upper_only, row_begin, row_end, d_c, d_cnt);
    mark(EV_NUMERIC); mark(EV_POST);
    unsigned long long* hc = reinterpret_cast<unsigned long long*>(static_cast<char*>(g.h_small) + 512);
    cudaError_t e = cudaSuccess;
    if (!rc) e = cudaMemcpyAsync(hc, d_cnt, 16, cudaMemcpyDeviceToHost, g.stream);
    mark(EV_D2H);
    if (!rc && e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    dfree(d_cnt);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "triple kernel", e);
    g.stats.products = (int64_t)(hc[0] + hc[1]);
    return SPGEMM_B200_OK;
}

int spgemm_b200_triple(int n, int k, const int32_t* h_indptr, const int32_t* h_indices, const double* h_values,
                       const int32_t* q_indptr, const int32_t* q_indices, const double* q_values, int mode,
                       double* c_host) {
    if (mode < 0 || mode > 2) return fail(SPGEMM_B200_ERR_ARG, "triple: bad mode");
    if (!c_host && n > 0) return fail(SPGEMM_B200_ERR_ARG, "triple: null output");
    int rc;
    if ((rc = check_csr_args(n, k, h_indptr, "triple: bad H"))) return rc;
    if ((rc = check_csr_args(k, k, q_indptr, "triple: bad Q"))) return rc;
    ENTER_DEFAULT();
    Ctx& g = cx();
    begin_call();
    spgemm_b200_mat *h = nullptr, *q = nullptr;
    double* d_c = nullptr;
    unsigned long long* d_cnt = nullptr;
    ZeroFill zero;                                 // upper mode: the host zeroes the lower triangle meanwhile
    if (mode == SPGEMM_B200_TRIPLE_UPPER) zero.start(c_host, n, 0, 1);
    auto done = [&](int code) {
        zero.join();
        dfree(d_c); dfree(d_cnt);
        mat_release(h); mat_release(q);
        return code;
    };
    {
        NvtxRange nv("spgemm_b200:h2d");
        if ((rc = upload(n, k, h_indptr, h_indices, h_values, &h))) return done(rc);
        if ((rc = upload(k, k, q_indptr, q_indices, q_values, &q))) return done(rc);
        mark(EV_H2D);
    }
    if ((rc = ensure_checked(h, q))) return done(rc);
    const size_t elems = (size_t)n * (size_t)n;
    if ((rc = dalloc(&d_c, elems)) || (rc = dalloc(&d_cnt, 4))) return done(rc);
    const bool upper = mode != SPGEMM_B200_TRIPLE_REF_FULL;
    if ((rc = triple_rows(h, q, nullptr, upper, 0, n, d_c, d_cnt))) return done(rc);
    mark(EV_NUMERIC);
    cudaError_t e = cudaSuccess;
    if (mode == SPGEMM_B200_TRIPLE_REF_FULL) e = launch_symmetrize(lctx(), d_c, n);
    else if (mode == SPGEMM_B200_TRIPLE_MIRROR) e = launch_mirror(lctx(), d_c, n);
    if (e != cudaSuccess) return done(fail(SPGEMM_B200_ERR_CUDA, "triple post kernel", e));
    mark(EV_POST);
    unsigned long long* hc = reinterpret_cast<unsigned long long*>(static_cast<char*>(g.h_small) + 512);
    e = cudaMemcpyAsync(hc, d_cnt, 16, cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess && elems) {
        NvtxRange nv("spgemm_b200:d2h");
        if (mode == SPGEMM_B200_TRIPLE_UPPER) e = d2h_upper_rows(d_c, n, 0, n, c_host);
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
    if (!d_c || n < 0) return fail(SPGEMM_B200_ERR_ARG, "mirror_dev: bad argument");
    ENTER_DEFAULT();
    CU(launch_mirror(lctx(), d_c, n));
    return SPGEMM_B200_OK;
}

int spgemm_b200_symmetrize_dev(double* d_c, int n) {
    if (!d_c || n < 0) return fail(SPGEMM_B200_ERR_ARG, "symmetrize_dev: bad argument");
    ENTER_DEFAULT();
    CU(launch_symmetrize(lctx(), d_c, n));
    return SPGEMM_B200_OK;
}

// ---- raw device buffers -----------------------------------------------------------------------------------
void* spgemm_b200_device_alloc(size_t bytes) {
    Ctx* c = default_ctx();
    if (!c) return nullptr;
    CallGuard guard(c);
    void* p = nullptr;
    cudaError_t e = cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, c->pool, c->stream);
    if (e != cudaSuccess) { fail(SPGEMM_B200_ERR_CUDA, "device_alloc", e); return nullptr; }
    return p;
}
void spgemm_b200_device_free(void* d_ptr) {
    if (!d_ptr) return;
    Ctx* c = g_default >= 0 ? live_ctx(g_default) : nullptr;
    if (!c) return;
    CallGuard guard(c);
    dfree(d_ptr);
}
int spgemm_b200_copy_to_host(void* host_dst, const void* d_src, size_t bytes) {
    if (bytes && (!host_dst || !d_src)) return fail(SPGEMM_B200_ERR_ARG, "copy_to_host: null pointer");
    ENTER_DEFAULT();
    if (bytes) CU(cudaMemcpyAsync(host_dst, d_src, bytes, cudaMemcpyDeviceToHost, cx().stream));
    CU(cudaStreamSynchronize(cx().stream));
    return SPGEMM_B200_OK;
}
int spgemm_b200_copy_to_device(void* d_dst, const void* host_src, size_t bytes) {
    if (bytes && (!d_dst || !host_src)) return fail(SPGEMM_B200_ERR_ARG, "copy_to_device: null pointer");
    ENTER_DEFAULT();
    if (bytes) CU(cudaMemcpyAsync(d_dst, host_src, bytes, cudaMemcpyHostToDevice, cx().stream));
    CU(cudaStreamSynchronize(cx().stream));
    return SPGEMM_B200_OK;
}

int spgemm_b200_copy_upper_to_host(double* host_dst, const double* d_src, int n) {
    if (n > 0 && (!host_dst || !d_src)) return fail(SPGEMM_B200_ERR_ARG, "copy_upper_to_host: null pointer");
    ENTER_DEFAULT();
    cx().stats.bytes_d2h = 0;
    ZeroFill zero;
    zero.start(host_dst, n, 0, 1);
    cudaError_t e = d2h_upper_rows(d_src, n, 0, n, host_dst);
    if (e == cudaSuccess) e = cudaStreamSynchronize(cx().stream);
    zero.join();
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "copy_upper_to_host", e);
    return SPGEMM_B200_OK;
}

int spgemm_b200_copy_on_device(void* d_dst, const void* d_src, size_t bytes) {
    if (bytes && (!d_dst || !d_src)) return fail(SPGEMM_B200_ERR_ARG, "copy_on_device: null pointer");
    ENTER_DEFAULT();
    if (bytes) CU(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, cx().stream));
    return SPGEMM_B200_OK;
}

// ---- peer memory ---------------------------------------------------------------------------------------------
void* spgemm_b200_shared_alloc(size_t bytes) {
    Ctx* c = default_ctx();
    if (!c) return nullptr;
    CallGuard guard(c);
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { fail(SPGEMM_B200_ERR_CUDA, "shared_alloc", e); return nullptr; }
    return p;
}
void spgemm_b200_shared_free(void* d_ptr) {
    if (!d_ptr) return;
    Ctx* c = g_default >= 0 ? live_ctx(g_default) : nullptr;
    if (!c) return;
    CallGuard guard(c);
    cudaFree(d_ptr);
}
int spgemm_b200_ipc_export(const void* d_ptr, unsigned char* handle) {
    if (!d_ptr || !handle) return fail(SPGEMM_B200_ERR_ARG, "ipc_export: null argument");
    ENTER_DEFAULT();
    static_assert(sizeof(cudaIpcMemHandle_t) == SPGEMM_B200_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)));
    memcpy(handle, &h, sizeof h);
    return SPGEMM_B200_OK;
}
int spgemm_b200_ipc_open(const unsigned char* handle, void** d_ptr) {
    if (!handle || !d_ptr) return fail(SPGEMM_B200_ERR_ARG, "ipc_open: null argument");
    ENTER_DEFAULT();
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SPGEMM_B200_OK;
}
int spgemm_b200_ipc_close(void* d_ptr) {
    if (!d_ptr) return SPGEMM_B200_OK;
    ENTER_DEFAULT();
    CU(cudaIpcCloseMemHandle(d_ptr));
    return SPGEMM_B200_OK;
}

// ---- stopwatch / L2 flush ------------------------------------------------------------------------------------
static const size_t kFlushBytes = (size_t)512 << 20;

int spgemm_b200_timer_start(void) {
    ENTER_DEFAULT();
    CU(cudaEventRecord(cx().t_ev0, cx().stream));
    return SPGEMM_B200_OK;
}
int spgemm_b200_timer_stop(double* ms) {
    ENTER_DEFAULT();
    CU(cudaEventRecord(cx().t_ev1, cx().stream));
    CU(cudaEventSynchronize(cx().t_ev1));
    float f = 0.f;
    CU(cudaEventElapsedTime(&f, cx().t_ev0, cx().t_ev1));
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
    ENTER_DEFAULT();
    Ctx& g = cx();
    if (!g.flush_buf) CU(cudaMalloc(&g.flush_buf, kFlushBytes + 256));
    CU(cudaMemsetAsync(g.flush_buf, 0x5a, kFlushBytes, g.stream));
    const char* base = static_cast<const char*>(g.flush_buf);
    k_flush_read<<<g.sm_count * 8, 256, 0, g.stream>>>(reinterpret_cast<const int4*>(base + kFlushBytes / 2),
                                                       kFlushBytes / 2 / sizeof(int4),
                                                       reinterpret_cast<int*>(const_cast<char*>(base) + kFlushBytes));
    CU(cudaGetLastError());
    return SPGEMM_B200_OK;
}

// ---- row costs / partition ------------------------------------------------------------------------------
int spgemm_b200_row_costs(const spgemm_b200_mat* a, const spgemm_b200_mat* b, const spgemm_b200_mat* q, int upper_only,
                          int dense_cols, int64_t* d_costs, int64_t* total_host) {
    if (!a || !b) return fail(SPGEMM_B200_ERR_ARG, "row_costs: null matrix");
    ENTER_DEVICE(a->device);
    Ctx& g = cx();
    int rc;
    if ((rc = ensure_checked(const_cast<spgemm_b200_mat*>(a), const_cast<spgemm_b200_mat*>(b)))) return rc;
    if (q && (rc = ensure_checked(const_cast<spgemm_b200_mat*>(q)))) return rc;
    const int m = a->rows;
    int64_t* costs = d_costs;
    if (!costs && (rc = dalloc(&costs, (size_t)m))) return rc;
    rc = row_costs_impl(a, b, q, upper_only, dense_cols, costs);
    cudaError_t e = cudaSuccess;
    if (!rc && total_host) {
        std::vector<int64_t> hc((size_t)m);          // total on the host (setup path, not timed)
        e = cudaMemcpyAsync(hc.data(), costs, (size_t)m * 8, cudaMemcpyDeviceToHost, g.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
        int64_t t = 0;
        for (int64_t v : hc) t += v;
        *total_host = t;
    }
    if (!d_costs) dfree(costs);
    if (rc) return rc;
    if (e != cudaSuccess) return fail(SPGEMM_B200_ERR_CUDA, "row_costs", e);
    return SPGEMM_B200_OK;
}

int spgemm_b200_partition_tail(const int64_t* d_costs, const int32_t* indptr_host, double tail_coeff, int rows, int parts,
                               int32_t* bounds_host) {
    if (!d_costs || !indptr_host || !bounds_host || rows < 0 || parts <= 0)
        return fail(SPGEMM_B200_ERR_ARG, "partition_tail: bad argument");
    ENTER_DEFAULT();
    std::vector<int64_t> c((size_t)rows), tail((size_t)rows + 1);
    if (rows) {
        CU(cudaMemcpyAsync(c.data(), d_costs, (size_t)rows * 8, cudaMemcpyDeviceToHost, cx().stream));
        CU(cudaStreamSynchronize(cx().stream));
    }
    for (int i = 0; i <= rows; ++i) tail[i] = (int64_t)indptr_host[rows] - indptr_host[i];
    partition_costs_tail(c.data(), tail.data(), tail_coeff, rows, parts, bounds_host);
    return SPGEMM_B200_OK;
}

int spgemm_b200_partition(const int64_t* d_costs, int rows, int parts, int32_t* bounds_host) {
    if (!d_costs || !bounds_host || rows < 0 || parts <= 0) return fail(SPGEMM_B200_ERR_ARG, "partition: bad argument");
    ENTER_DEFAULT();
    std::vector<int64_t> c((size_t)rows);
    if (rows) {
        CU(cudaMemcpyAsync(c.data(), d_costs, (size_t)rows * 8, cudaMemcpyDeviceToHost, cx().stream));
        CU(cudaStreamSynchronize(cx().stream));
    }
    partition_costs(c.data(), rows, parts, bounds_host);
    return SPGEMM_B200_OK;
}

}  // extern "C"
