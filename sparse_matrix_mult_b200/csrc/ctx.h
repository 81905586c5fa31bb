// ctx.h -- host-side internals shared by api.cu (single-device C ABI) and multi.cu (single-process multi-GPU
// driver): the per-device context, the per-call guard, handle layouts and the implementation functions the entry
// points are built from.  Not part of the ABI.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/spgemm_b200.h"
#include "internal.h"

// ---------------------------------------------------------------------------------------------------
// handles
struct spgemm_b200_mat {
    int rows, cols;
    int64_t nnz;
    int32_t* ptr;
    int32_t* idx;
    double* val;
    bool owns;
    int device;                // ordinal of the device the arrays live on
    int32_t* d_flags;          // device int32[8]: [0] rows sorted ascending (read by the kernels), [1..5] check scratch
    bool checked;              // the validation pass ran and its verdict is in `sorted` / `valid` / `runs`
    bool sorted, valid;
    bool runs;                 // every row is one run of consecutive ascending columns (banded matrix)
    bool desc_sorted;          // rows sorted by DESCENDING column (set by the transpose)
    spgemm_b200_mat* shadow;   // row-sorted copy of a borrowed (owns == false) unsorted matrix, built on demand
    bool cache_panels;         // keep the paneled transpose the triple product builds from this matrix (as H)
    void* panel_cache;         // the kept transpose (api.cu: PanelCache), or null
};
struct spgemm_b200_result {
    int rows, cols;
    int64_t nnz;
    int64_t* d_ptr;
    int32_t* d_idx;
    double* d_val;
    int device;
};

namespace sbh {

enum { EV_START = 0, EV_H2D, EV_ANALYSIS, EV_SYMBOLIC, EV_NUMERIC, EV_POST, EV_D2H, EV_COUNT };

constexpr int kMaxDevices = 16;

// One per CUDA device, created on first use.  Every public entry point holds `mu` for its whole duration (the
// library keeps per-call state here: stats, event mask, the pinned staging block), so concurrent callers on one
// device are serialised and callers on different devices run in parallel.
struct Ctx {
    std::recursive_mutex mu;
    bool ready = false;
    int device = 0;
    int sm_count = 0;
    cudaMemPool_t pool = nullptr;                 // private stream-ordered pool (not the device's default pool)
    cudaStream_t own_stream = nullptr, stream = nullptr;
    cudaStream_t aux[3] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[3] = {};
    cudaEvent_t ev[EV_COUNT] = {};
    cudaEvent_t t_ev0 = nullptr, t_ev1 = nullptr; // stopwatch (timer_start / timer_stop)
    bool ev_pending = false;
    unsigned ev_mask = 0;                         // which events were recorded by the current call
    spgemm_b200_stats stats = {};
    int launches = 0;
    void* h_small = nullptr;                      // 4 KB pinned staging for counters
    void* flush_buf = nullptr;
};

// Locks the context, makes its device current (restoring the caller's device on exit) and publishes it as this
// thread's current context for the helpers below.  Nestable.
class CallGuard {
public:
    explicit CallGuard(Ctx* c);
    ~CallGuard();
    CallGuard(const CallGuard&) = delete;
    CallGuard& operator=(const CallGuard&) = delete;
private:
    Ctx* ctx_;
    Ctx* prev_ctx_;
    int prev_dev_;
};

Ctx& cx();                                        // this thread's current context (inside a CallGuard)
Ctx* default_ctx();                               // initialised default-device context, or nullptr after fail()
Ctx* device_ctx(int device);                      // initialised context of `device`, or nullptr after fail()
Ctx* live_ctx(int device);                        // context of `device` if it exists already, else nullptr

int fail(int code, const char* what, cudaError_t e = cudaSuccess);

#define CU(call)                                                               \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess) return sbh::fail(SPGEMM_B200_ERR_CUDA, #call, e__); \
    } while (0)

struct NvtxRange {                                // one range per phase of a call (visible in nsys / ncu timelines)
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

template <typename T>
int dalloc(T** p, size_t count) {
    *p = nullptr;
    Ctx& g = cx();
    CU(cudaMallocFromPoolAsync((void**)p, (count ? count : 1) * sizeof(T), g.pool, g.stream));
    return SPGEMM_B200_OK;
}
void dfree(void* p);

sb::LaunchCtx lctx();
inline sb::Csr view(const spgemm_b200_mat* m) { return sb::Csr{m->ptr, m->idx, m->val, m->rows, m->cols}; }
inline int64_t csr_bytes(int64_t rows, int64_t nnz) { return 12 * nnz + 4 * (rows + 1); }

void mark(int ev);
void begin_call(bool lean = false);
void finish_stats();

// ---- building blocks (all run on cx()) ----------------------------------------------------------------
int upload(int rows, int cols, const int32_t* ptr, const int32_t* idx, const double* val, spgemm_b200_mat** out);
int alloc_mat(int rows, int cols, int64_t nnz, spgemm_b200_mat** out);   // device arrays only, nothing copied
void mat_release(spgemm_b200_mat* m);             // frees device arrays on cx() and deletes the handle
void result_release(spgemm_b200_result* r);
int transpose_impl(const spgemm_b200_mat* x, spgemm_b200_mat** out, bool sort_desc);
// validation + sortedness of up to two matrices with ONE host synchronisation; ERR_ARG when an index is out of
// range or an indptr is not monotone
int ensure_checked(spgemm_b200_mat* m1, spgemm_b200_mat* m2 = nullptr);
// the matrix itself when its rows are sorted, else a row-sorted version (in place when the library owns the
// arrays, a cached shadow copy otherwise)
int sorted_view(spgemm_b200_mat* m, spgemm_b200_mat** out);
int csr_impl(spgemm_b200_mat* a, spgemm_b200_mat* b, int upper_only, int r0, int r1, spgemm_b200_result** out);
int dense_rows(spgemm_b200_mat* a, spgemm_b200_mat* b, int upper_only, int r0, int r1, double* d_c);
// rows [r0, r1) of H Q H^T into d_c.  Builds the paneled transpose of H it needs (or uses `ht`, a plain transpose of
// H, when one panel is enough); records EV_ANALYSIS / EV_SYMBOLIC before the kernel.  d_cnt: device u64[4].
int triple_rows(const spgemm_b200_mat* h, const spgemm_b200_mat* q, const spgemm_b200_mat* ht, int upper_only, int r0,
                int r1, double* d_c, unsigned long long* d_cnt);
int row_costs_impl(const spgemm_b200_mat* a, const spgemm_b200_mat* b, const spgemm_b200_mat* q, int upper_only,
                   int dense_cols, int64_t* d_costs);
void partition_costs(const int64_t* costs, int rows, int parts, int32_t* bounds);
// the same with a fixed cost per block that depends on its first row: tail_coeff * tail[first row]
void partition_costs_tail(const int64_t* costs, const int64_t* tail, double tail_coeff, int rows, int parts, int32_t* bounds);
constexpr double kTripleTailCoeff = 8.6;      // cost units per entry of H a rank has to transpose (k_triple_costs units)
// rows [r0, r1) of an n-column result whose entries left of the diagonal are zero: columns [block start, n) of
// each fixed row block cross PCIe (asynchronous, on the context's stream).  d_c holds the rows [r0, r1) only.
cudaError_t d2h_upper_rows(const double* d_c, int n, int r0, int r1, double* c_host);
// host threads zeroing share `part` of `nparts` of the rectangles left of those blocks (the lower triangle of the
// n x n host matrix); started at the beginning of a call, joined before it returns
class ZeroFill {
public:
    void start(double* c_host, int n, int part, int nparts);
    void join();
    ~ZeroFill() { join(); }
private:
    std::vector<std::thread> threads_;
};

// pinned host cache (process-wide)
void* host_cache_alloc(size_t bytes);
void host_cache_free(void* p);
void host_cache_clear();

}  // namespace sbh
