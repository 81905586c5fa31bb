// Micro-benchmark: what write-only bandwidth can a kernel reach on this B200?  (ceiling for the dense kernels)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fill_bw fill_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

__global__ void fill_cs_v2(double* p, size_t n2) {   // n2 = number of double2
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.cs.v2.f64 [%0], {%1, %1};" ::"l"(p + 2 * i), "d"(0.0) : "memory");
}
__global__ void fill_wb_v2(double* p, size_t n2) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.v2.f64 [%0], {%1, %1};" ::"l"(p + 2 * i), "d"(0.0) : "memory");
}
__global__ void fill_v4(double* p, size_t n4) {      // 256-bit stores (sm_100+)
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p + 4 * i), "d"(0.0) : "memory");
}
// one block per row of `cols` doubles (the access pattern of k_dense_rows_red)
__global__ void fill_rows(double* p, int rows, int cols) {
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        double* row = p + (size_t)r * cols;
        for (int i = threadIdx.x; i < cols / 2; i += blockDim.x)
            asm volatile("st.global.cs.v2.f64 [%0], {%1, %1};" ::"l"(row + 2 * i), "d"(0.0) : "memory");
    }
}
__global__ void fill_rows_v4(double* p, int rows, int cols) {
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        double* row = p + (size_t)r * cols;
        for (int i = threadIdx.x; i < cols / 4; i += blockDim.x)
            asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(row + 4 * i), "d"(0.0) : "memory");
    }
}
// bulk (TMA) store of a zeroed shared-memory buffer, `chunk` bytes per copy
__global__ void fill_bulk(double* p, size_t total_bytes, int chunk) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x) reinterpret_cast<int4*>(sm)[i] = make_int4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned saddr = (unsigned)__cvta_generic_to_shared(sm);
        size_t nchunks = total_bytes / chunk;
        int inflight = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
            char* dst = reinterpret_cast<char*>(p) + c * (size_t)chunk;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <class F> float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    const int rows = 20000, cols = 20000;
    const size_t n = (size_t)rows * cols, bytes = n * 8;
    double* p; CK(cudaMalloc(&p, bytes));
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    auto rep = [&](const char* name, float ms) { printf("%-34s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / 1e6 / ms); };
    rep("cudaMemsetAsync", timeit([&] { cudaMemsetAsync(p, 0, bytes); }));
    for (int mult : {2, 4, 8, 16}) {
        char nm[64];
        snprintf(nm, 64, "fill_cs_v2 grid=%dxSM x256", mult);   rep(nm, timeit([&] { fill_cs_v2<<<sms * mult, 256>>>(p, n / 2); }));
        snprintf(nm, 64, "fill_wb_v2 grid=%dxSM x256", mult);   rep(nm, timeit([&] { fill_wb_v2<<<sms * mult, 256>>>(p, n / 2); }));
        snprintf(nm, 64, "fill_v4    grid=%dxSM x256", mult);   rep(nm, timeit([&] { fill_v4<<<sms * mult, 256>>>(p, n / 4); }));
        snprintf(nm, 64, "fill_rows  grid=%dxSM x256", mult);   rep(nm, timeit([&] { fill_rows<<<sms * mult, 256>>>(p, rows, cols); }));
        snprintf(nm, 64, "fill_rows_v4 grid=%dxSM x256", mult); rep(nm, timeit([&] { fill_rows_v4<<<sms * mult, 256>>>(p, rows, cols); }));
    }
    rep("fill_rows grid=rows x256", timeit([&] { fill_rows<<<rows, 256>>>(p, rows, cols); }));
    rep("fill_rows grid=rows x512", timeit([&] { fill_rows<<<rows, 512>>>(p, rows, cols); }));
    for (int chunk : {16384, 32768, 65536}) {
        cudaFuncSetAttribute(fill_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk);
        for (int mult : {1, 2, 4}) {
            char nm[64]; snprintf(nm, 64, "fill_bulk chunk=%dK grid=%dxSM", chunk / 1024, mult);
            rep(nm, timeit([&] { fill_bulk<<<sms * mult, 128, chunk>>>(p, bytes, chunk); }));
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
