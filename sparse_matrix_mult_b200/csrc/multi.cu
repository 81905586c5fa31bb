// multi.cu -- single-process multi-GPU driver behind the drop-in API (VERDICT r1 "next" #6/#7).
//
// The reference sizes its OpenMP team INSIDE the call (omp_get_max_threads(), src/sparse_sparse_sparse.cpp:188-197)
// and hands each thread a contiguous row range from limits() (src/workdivision.cpp:16-89).  The equivalent here:
// one host thread per GPU inside the call, each with its own context (stream, pool), and a contiguous row range
// from the flop-balanced partition.  The path shards by output rows with no exchange between shards, so there
// is no collective in the data path:
//     every GPU : H2D of 1/N of every operand array over ITS OWN PCIe link, then the other N-1 parts from its peers
//                 over NVLink (peer copies: a scatter + all-gather, so the host is read once instead of N times;
//                 whole-operand uploads per GPU when the devices cannot reach each other) -> checks (-> H^T)
//     GPU 0     : per-row cost pass -> partition (published to the others through a barrier)
//     every GPU : its row block with the same kernels as the single-GPU path
//                 -> D2H of its block straight into its rows of the caller's result (N PCIe links at once;
//                    the symmetric dense modes send upper trapezoids only, host threads zero the rest)
// The one-process-per-GPU path (torch.distributed + NCCL broadcast / gather to rank 0) is distributed.py.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ctx.h"

using namespace sb;
using namespace sbh;

struct spgemm_b200_multi_result {
    int n_gpus, rows, cols;
    int64_t nnz;
    std::vector<int32_t> bounds;                     // n_gpus + 1 row bounds
    std::vector<spgemm_b200_result*> parts;          // one device-resident block per GPU
};

namespace {

class Barrier {
public:
    explicit Barrier(int n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu_);
        const int gen = gen_;
        if (++count_ == n_) {
            count_ = 0;
            ++gen_;
            cv_.notify_all();
        } else {
            cv_.wait(lk, [&] { return gen != gen_; });
        }
    }
private:
    std::mutex mu_;
    std::condition_variable cv_;
    int n_, count_ = 0, gen_ = 0;
};

enum Kind { K_DENSE, K_TRIPLE, K_CSR };

struct Operand {
    int rows, cols;
    const int32_t *ptr, *idx;
    const double* val;
};

struct Shard {                    // what a worker publishes for the peer all-gather of the operands
    void* base[2][3] = {};        // [operand][ptr, idx, val] device arrays
    cudaEvent_t ready = nullptr;  // its own slices have landed
};

struct Job {
    Kind kind;
    int n_gpus;
    bool peers = false;           // every GPU can read every other GPU's memory
    std::vector<Shard> shards;
    Operand a, b;                 // dense/csr: A, B;  triple: H, Q
    int upper_only;
    double* c_host = nullptr;     // dense / triple
    spgemm_b200_multi_result* res = nullptr;
    std::vector<int32_t> bounds;
    Barrier bar;
    std::mutex mu;
    int status = SPGEMM_B200_OK;
    std::string err;
    std::vector<spgemm_b200_stats> stats;
    Job(int n) : n_gpus(n), shards(n), bounds(n + 1, 0), bar(n), stats(n) {}
    void set_error(int code) {
        std::lock_guard<std::mutex> lk(mu);
        if (status == SPGEMM_B200_OK) { status = code; err = spgemm_b200_last_error(); }
    }
    bool failed() {
        std::lock_guard<std::mutex> lk(mu);
        return status != SPGEMM_B200_OK;
    }
};

std::mutex g_multi_mu;                         // one multi-GPU call at a time
std::vector<int32_t> g_last_bounds;
std::vector<spgemm_b200_stats> g_last_stats;

// Peer access between the first n devices (both the driver-level switch and the access list of the private pools),
// set up once per process.  False when some pair cannot reach each other.
bool ensure_peers(int n) {
    static std::mutex mu;
    static int have = 0;          // peers are set up among devices [0, have)
    static bool ok = true;
    std::lock_guard<std::mutex> lk(mu);
    if (n <= have) return ok;
    for (int d = 0; d < n && ok; ++d) {
        Ctx* c = device_ctx(d);
        if (!c) { ok = false; break; }
        CallGuard guard(c);
        for (int p = 0; p < n; ++p) {
            if (p == d) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, d, p) != cudaSuccess || !can) { ok = false; break; }
            cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) { cudaGetLastError(); ok = false; break; }
            cudaMemAccessDesc desc = {};
            desc.location.type = cudaMemLocationTypeDevice;
            desc.location.id = p;
            desc.flags = cudaMemAccessFlagsProtReadWrite;
            if (cudaMemPoolSetAccess(c->pool, &desc, 1) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
        }
    }
    have = n;
    return ok;
}

// This worker's 1/N slice of every operand array, host -> its own device (scatter half of the operand broadcast).
int upload_slices(Job* job, int d, const Operand& op, int which, spgemm_b200_mat** out) {
    Ctx& g = cx();
    const int64_t nnz = op.rows > 0 ? (int64_t)op.ptr[op.rows] - op.ptr[0] : 0;
    if (nnz < 0 || (op.rows > 0 && op.ptr[0] != 0)) return fail(SPGEMM_B200_ERR_ARG, "multi: bad indptr");
    if (nnz > 0 && (!op.idx || !op.val)) return fail(SPGEMM_B200_ERR_ARG, "multi: null indices/values with nnz > 0");
    spgemm_b200_mat* m = nullptr;
    int rc = alloc_mat(op.rows, op.cols, nnz, &m);
    if (rc) return rc;
    const void* host[3] = {op.ptr, op.idx, op.val};
    void* dev[3] = {m->ptr, m->idx, m->val};
    const size_t count[3] = {(size_t)op.rows + 1, (size_t)nnz, (size_t)nnz}, width[3] = {4, 4, 8};
    const int N = job->n_gpus;
    for (int k = 0; k < 3; ++k) {
        const size_t lo = count[k] * d / N, hi = count[k] * (d + 1) / N;
        job->shards[d].base[which][k] = dev[k];
        if (hi > lo) {
            cudaError_t e = cudaMemcpyAsync((char*)dev[k] + lo * width[k], (const char*)host[k] + lo * width[k],
                                            (hi - lo) * width[k], cudaMemcpyHostToDevice, g.stream);
            if (e != cudaSuccess) { mat_release(m); return fail(SPGEMM_B200_ERR_CUDA, "multi: slice upload", e); }
            g.stats.bytes_h2d += (int64_t)((hi - lo) * width[k]);
        }
    }
    *out = m;
    return SPGEMM_B200_OK;
}

// The other N-1 slices of an operand from the peers that uploaded them (all-gather half, NVLink).
int gather_slices(Job* job, int d, const spgemm_b200_mat* m, int which) {
    Ctx& g = cx();
    void* dev[3] = {m->ptr, m->idx, m->val};
    const size_t count[3] = {(size_t)m->rows + 1, (size_t)m->nnz, (size_t)m->nnz}, width[3] = {4, 4, 8};
    const int N = job->n_gpus;
    for (int step = 1; step < N; ++step) {
        const int s = (d + step) % N;                 // every worker starts at a different peer
        for (int k = 0; k < 3; ++k) {
            const size_t lo = count[k] * s / N, hi = count[k] * (s + 1) / N;
            if (hi <= lo) continue;
            CU(cudaMemcpyPeerAsync((char*)dev[k] + lo * width[k], d, (const char*)job->shards[s].base[which][k] + lo * width[k],
                                   s, (hi - lo) * width[k], g.stream));
        }
    }
    return SPGEMM_B200_OK;
}

// Every worker reaches every barrier whatever happens (a worker that failed just stops doing work).
void worker(Job* job, int d) {
    int rc = SPGEMM_B200_OK;
    Ctx* c = device_ctx(d);
    if (!c) {
        job->set_error(SPGEMM_B200_ERR_CUDA);
        if (job->peers) job->bar.wait();
        job->bar.wait(); job->bar.wait();
        return;
    }
    CallGuard guard(c);
    Ctx& g = cx();
    begin_call();
    // symmetric dense result: every worker zeroes its share of the lower triangle on host threads, from the start
    ZeroFill zero;
    {
        const bool square = job->kind == K_TRIPLE || job->a.rows == job->b.cols;
        if (job->kind != K_CSR && job->upper_only && square)
            zero.start(job->c_host, job->kind == K_TRIPLE ? job->a.rows : job->b.cols, d, job->n_gpus);
    }
    spgemm_b200_mat *a = nullptr, *b = nullptr, *ht = nullptr;
    double* d_c = nullptr;
    unsigned long long* d_cnt = nullptr;
    int64_t* d_costs = nullptr;
    const bool same = job->kind == K_CSR && job->a.ptr == job->b.ptr && job->a.idx == job->b.idx &&
                      job->a.val == job->b.val && job->a.rows == job->b.rows && job->a.cols == job->b.cols;
    const int m = job->a.rows;
    // ---- phase 1: operands in, checks, H^T, (GPU 0) partition ----
    if (job->peers) {
        NvtxRange nv("spgemm_b200:multi:scatter+allgather");
        rc = upload_slices(job, d, job->a, 0, &a);
        if (!rc) {
            if (same) b = a;
            else rc = upload_slices(job, d, job->b, 1, &b);
        }
        cudaError_t e = cudaSuccess;
        if (!rc) e = cudaEventCreateWithFlags(&job->shards[d].ready, cudaEventDisableTiming);
        if (!rc && e == cudaSuccess) e = cudaEventRecord(job->shards[d].ready, g.stream);
        if (!rc && e != cudaSuccess) rc = fail(SPGEMM_B200_ERR_CUDA, "multi: slice event", e);
        if (rc) job->set_error(rc);
        job->bar.wait();                                   // every worker's arrays and event are published
        if (!job->failed()) {
            for (int s = 0; s < job->n_gpus && !rc; ++s)
                if (s != d && cudaStreamWaitEvent(g.stream, job->shards[s].ready, 0) != cudaSuccess)
                    rc = fail(SPGEMM_B200_ERR_CUDA, "multi: wait for a peer's slices");
            if (!rc) rc = gather_slices(job, d, a, 0);
            if (!rc && b != a) rc = gather_slices(job, d, b, 1);
        } else if (!rc) {
            rc = SPGEMM_B200_ERR_STATE;                   // another worker failed: skip the work, keep the barriers
        }
        mark(EV_H2D);
    } else {
        NvtxRange nv("spgemm_b200:multi:h2d");
        rc = upload(job->a.rows, job->a.cols, job->a.ptr, job->a.idx, job->a.val, &a);
        if (!rc) {
            if (same) b = a;
            else rc = upload(job->b.rows, job->b.cols, job->b.ptr, job->b.idx, job->b.val, &b);
        }
        mark(EV_H2D);
    }
    if (!rc) rc = ensure_checked(a, b);
    if (!rc && d == 0 && job->kind == K_TRIPLE) rc = transpose_impl(a, &ht, false);   // row lengths of H^T for the costs
    if (!rc && d == 0) {
        std::vector<int64_t> costs((size_t)m);
        rc = dalloc(&d_costs, (size_t)m);
        if (!rc) rc = row_costs_impl(a, job->kind == K_TRIPLE ? ht : b, job->kind == K_TRIPLE ? b : nullptr,
                                     job->upper_only, job->kind == K_DENSE ? job->b.cols : 0, d_costs);
        if (!rc && m > 0) {
            cudaError_t e = cudaMemcpyAsync(costs.data(), d_costs, (size_t)m * 8, cudaMemcpyDeviceToHost, g.stream);
            if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
            if (e != cudaSuccess) rc = fail(SPGEMM_B200_ERR_CUDA, "multi: row costs copy", e);
        }
        if (!rc) {
            if (job->kind == K_TRIPLE) {                      // the rank that starts at row r transposes rows r.. of H
                std::vector<int64_t> tail((size_t)m + 1);
                for (int i = 0; i <= m; ++i) tail[i] = (int64_t)job->a.ptr[m] - job->a.ptr[i];
                partition_costs_tail(costs.data(), tail.data(), kTripleTailCoeff, m, job->n_gpus, job->bounds.data());
            } else {
                partition_costs(costs.data(), m, job->n_gpus, job->bounds.data());
            }
        }
        dfree(d_costs);
    }
    mark(EV_ANALYSIS);
    if (rc) job->set_error(rc);
    job->bar.wait();
    // ---- phase 2: this GPU's row block ----
    if (!job->failed()) {
        const int r0 = job->bounds[d], r1 = job->bounds[d + 1];
        const int n = job->kind == K_TRIPLE ? job->a.rows : job->b.cols;
        if (job->kind == K_CSR) {
            spgemm_b200_result* part = nullptr;
            rc = csr_impl(a, b, job->upper_only, r0, r1, &part);
            if (!rc) job->res->parts[d] = part;
            mark(EV_POST); mark(EV_D2H);
        } else {
            const bool square = job->kind == K_TRIPLE || job->a.rows == job->b.cols;
            if (r1 > r0) {
                if (job->kind != K_TRIPLE) mark(EV_SYMBOLIC);
                rc = dalloc(&d_c, (size_t)(r1 - r0) * (size_t)n);
                if (!rc && job->kind == K_TRIPLE) {
                    rc = dalloc(&d_cnt, 4);
                    if (!rc) rc = triple_rows(a, b, nullptr, job->upper_only, r0, r1, d_c, d_cnt);
                } else if (!rc) {
                    rc = dense_rows(a, b, job->upper_only, r0, r1, d_c);
                }
                mark(EV_NUMERIC); mark(EV_POST);
            }
            if (!rc) {
                NvtxRange nv("spgemm_b200:multi:d2h");
                cudaError_t e = cudaSuccess;
                if (job->upper_only && square) {
                    if (r1 > r0) e = d2h_upper_rows(d_c, n, r0, r1, job->c_host);
                } else if (r1 > r0) {
                    e = cudaMemcpyAsync(job->c_host + (size_t)r0 * n, d_c, (size_t)(r1 - r0) * n * 8,
                                        cudaMemcpyDeviceToHost, g.stream);
                    g.stats.bytes_d2h += (int64_t)(r1 - r0) * n * 8;
                }
                if (e != cudaSuccess) rc = fail(SPGEMM_B200_ERR_CUDA, "multi: result copy", e);
            }
            mark(EV_D2H);
        }
        if (!rc) {
            cudaError_t e = cudaStreamSynchronize(g.stream);
            if (e != cudaSuccess) rc = fail(SPGEMM_B200_ERR_CUDA, "multi: synchronize", e);
        }
        if (rc) job->set_error(rc);
    }
    cudaStreamSynchronize(g.stream);
    zero.join();
    finish_stats();
    job->stats[d] = g.stats;
    dfree(d_c); dfree(d_cnt);
    if (b != a) mat_release(b);
    mat_release(a);
    mat_release(ht);
    job->bar.wait();
    if (job->shards[d].ready) cudaEventDestroy(job->shards[d].ready);
}

int run(Job& job) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    const char* no_peer = getenv("SPGEMM_B200_MULTI_NO_PEER");     // experiments: whole-operand upload on every GPU
    job.peers = job.n_gpus > 1 && !(no_peer && atoi(no_peer)) && ensure_peers(job.n_gpus);
    std::vector<std::thread> threads;
    for (int d = 1; d < job.n_gpus; ++d) threads.emplace_back(worker, &job, d);
    worker(&job, 0);
    for (auto& t : threads) t.join();
    g_last_bounds = job.bounds;
    g_last_stats = job.stats;
    if (job.status != SPGEMM_B200_OK) return fail(job.status, job.err.c_str());
    return SPGEMM_B200_OK;
}

int check_gpus(int n_gpus) {
    if (n_gpus < 1 || n_gpus > kMaxDevices) return fail(SPGEMM_B200_ERR_ARG, "multi: n_gpus out of range");
    if (n_gpus > spgemm_b200_device_count()) return fail(SPGEMM_B200_ERR_ARG, "multi: more GPUs requested than visible");
    return SPGEMM_B200_OK;
}

}  // namespace

extern "C" {

int spgemm_b200_multi_dense(int n_gpus, int m, int k, int n, const int32_t* a_indptr, const int32_t* a_indices,
                            const double* a_values, const int32_t* b_indptr, const int32_t* b_indices,
                            const double* b_values, int upper_only, double* c_host) {
    int rc = check_gpus(n_gpus);
    if (rc) return rc;
    if (m < 0 || k < 0 || n < 0 || !a_indptr || !b_indptr) return fail(SPGEMM_B200_ERR_ARG, "multi_dense: bad operand");
    if (!c_host && (int64_t)m * n > 0) return fail(SPGEMM_B200_ERR_ARG, "multi_dense: null output");
    Job job(n_gpus);
    job.kind = K_DENSE;
    job.a = Operand{m, k, a_indptr, a_indices, a_values};
    job.b = Operand{k, n, b_indptr, b_indices, b_values};
    job.upper_only = upper_only;
    job.c_host = c_host;
    return run(job);
}

int spgemm_b200_multi_triple(int n_gpus, int n, int k, const int32_t* h_indptr, const int32_t* h_indices,
                             const double* h_values, const int32_t* q_indptr, const int32_t* q_indices,
                             const double* q_values, double* c_host) {
    int rc = check_gpus(n_gpus);
    if (rc) return rc;
    if (n < 0 || k < 0 || !h_indptr || !q_indptr) return fail(SPGEMM_B200_ERR_ARG, "multi_triple: bad operand");
    if (!c_host && n > 0) return fail(SPGEMM_B200_ERR_ARG, "multi_triple: null output");
    Job job(n_gpus);
    job.kind = K_TRIPLE;
    job.a = Operand{n, k, h_indptr, h_indices, h_values};
    job.b = Operand{k, k, q_indptr, q_indices, q_values};
    job.upper_only = 1;
    job.c_host = c_host;
    return run(job);
}

int spgemm_b200_multi_csr(int n_gpus, int m, int k, int n, const int32_t* a_indptr, const int32_t* a_indices,
                          const double* a_values, const int32_t* b_indptr, const int32_t* b_indices,
                          const double* b_values, int upper_only, spgemm_b200_multi_result** out) {
    int rc = check_gpus(n_gpus);
    if (rc) return rc;
    if (m < 0 || k < 0 || n < 0 || !a_indptr || !b_indptr || !out) return fail(SPGEMM_B200_ERR_ARG, "multi_csr: bad argument");
    spgemm_b200_multi_result* res = new spgemm_b200_multi_result{n_gpus, m, n, 0, {}, std::vector<spgemm_b200_result*>(n_gpus, nullptr)};
    Job job(n_gpus);
    job.kind = K_CSR;
    job.a = Operand{m, k, a_indptr, a_indices, a_values};
    job.b = Operand{k, n, b_indptr, b_indices, b_values};
    job.upper_only = upper_only;
    job.res = res;
    rc = run(job);
    res->bounds = job.bounds;
    if (rc) { spgemm_b200_multi_result_free(res); return rc; }
    for (auto* p : res->parts) res->nnz += p ? p->nnz : 0;
    *out = res;
    return SPGEMM_B200_OK;
}

int64_t spgemm_b200_multi_result_nnz(const spgemm_b200_multi_result* r) { return r ? r->nnz : -1; }

// Parallel copy-out: every GPU sends its indices/values over its own PCIe link to their final offsets in the
// caller's arrays; the row pointers are rebased on the host (the parallel replacement of the reference's serial
// stitch, src/sparse_sparse_sparse.cpp:265-291).
int spgemm_b200_multi_result_copy(const spgemm_b200_multi_result* r, void* indptr, int index64, int32_t* indices,
                                  double* values) {
    if (!r || !indptr) return fail(SPGEMM_B200_ERR_ARG, "multi_result_copy: null argument");
    if (r->nnz > 0 && (!indices || !values)) return fail(SPGEMM_B200_ERR_ARG, "multi_result_copy: null indices/values");
    if (!index64 && r->nnz > 0x7fffffffLL) return fail(SPGEMM_B200_ERR_OVERFLOW, "nnz(C) >= 2^31 needs index64");
    std::vector<int64_t> offs(r->n_gpus + 1, 0);
    for (int d = 0; d < r->n_gpus; ++d) offs[d + 1] = offs[d] + (r->parts[d] ? r->parts[d]->nnz : 0);
    std::mutex mu;
    int status = SPGEMM_B200_OK;
    std::string err;
    auto copy_part = [&](int d) {
        const spgemm_b200_result* p = r->parts[d];
        if (!p) return;
        Ctx* c = device_ctx(p->device);
        int rc = SPGEMM_B200_OK;
        if (!c) rc = SPGEMM_B200_ERR_CUDA;
        else {
            CallGuard guard(c);
            Ctx& g = cx();
            const int rows = p->rows;
            std::vector<int64_t> local((size_t)rows + 1);
            cudaError_t e = cudaMemcpyAsync(local.data(), p->d_ptr, ((size_t)rows + 1) * 8, cudaMemcpyDeviceToHost, g.stream);
            if (e == cudaSuccess && p->nnz > 0) {
                e = cudaMemcpyAsync(indices + offs[d], p->d_idx, (size_t)p->nnz * 4, cudaMemcpyDeviceToHost, g.stream);
                if (e == cudaSuccess)
                    e = cudaMemcpyAsync(values + offs[d], p->d_val, (size_t)p->nnz * 8, cudaMemcpyDeviceToHost, g.stream);
            }
            if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
            if (e != cudaSuccess) rc = fail(SPGEMM_B200_ERR_CUDA, "multi_result_copy", e);
            else {
                const int r0 = r->bounds[d];
                if (index64) {
                    int64_t* out = static_cast<int64_t*>(indptr);
                    for (int i = 0; i <= rows; ++i) out[r0 + i] = local[i] + offs[d];
                } else {
                    int32_t* out = static_cast<int32_t*>(indptr);
                    for (int i = 0; i <= rows; ++i) out[r0 + i] = (int32_t)(local[i] + offs[d]);
                }
            }
        }
        if (rc) {
            std::lock_guard<std::mutex> lk(mu);
            if (!status) { status = rc; err = spgemm_b200_last_error(); }
        }
    };
    std::vector<std::thread> threads;
    for (int d = 1; d < r->n_gpus; ++d) threads.emplace_back(copy_part, d);
    copy_part(0);
    for (auto& t : threads) t.join();
    if (status) return fail(status, err.c_str());
    return SPGEMM_B200_OK;
}

void spgemm_b200_multi_result_free(spgemm_b200_multi_result* r) {
    if (!r) return;
    for (auto* p : r->parts) spgemm_b200_result_free(p);
    delete r;
}

int spgemm_b200_multi_last_bounds(int32_t* bounds, int capacity) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    const int n = (int)g_last_bounds.size();
    for (int i = 0; i < n && i < capacity; ++i) bounds[i] = g_last_bounds[i];
    return n;
}

int spgemm_b200_multi_last_stats(int part, spgemm_b200_stats* out) {
    std::lock_guard<std::mutex> lk(g_multi_mu);
    if (!out || part < 0 || part >= (int)g_last_stats.size()) return fail(SPGEMM_B200_ERR_ARG, "multi_last_stats: bad argument");
    *out = g_last_stats[part];
    return SPGEMM_B200_OK;
}

}  // extern "C"
