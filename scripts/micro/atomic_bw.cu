// Micro-benchmark: float64 accumulate throughput on B200 -- L2 reductions (RED.ADD.F64) vs shared-memory
// atomics (CAS loop) vs plain shared RMW.  Informs the accumulator design of the SpGEMM kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_bw atomic_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
__device__ __forceinline__ unsigned rng(unsigned& s) { s = s * 1664525u + 1013904223u; return s; }

// each block adds `per_thread` values per thread at pseudo-random positions of ITS OWN slice of `span` doubles
__global__ void red_global(double* buf, size_t span, int per_thread, int shared_slice) {
    double* mine = buf + (shared_slice ? 0 : (size_t)blockIdx.x * span);
    unsigned s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
    for (int i = 0; i < per_thread; ++i) atomicAdd(mine + (rng(s) >> 7) % span, 1.0);
}
__global__ void atom_shared(double* out, int span, int per_thread) {
    extern __shared__ double acc[];
    for (int i = threadIdx.x; i < span; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    unsigned s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
    for (int i = 0; i < per_thread; ++i) atomicAdd(acc + (rng(s) >> 7) % span, 1.0);
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = acc[0];
}
__global__ void plain_shared(double* out, int span, int per_thread) {   // racy on purpose: cost of LDS+DADD+STS only
    extern __shared__ double acc[];
    for (int i = threadIdx.x; i < span; i += blockDim.x) acc[i] = 0;
    __syncthreads();
    unsigned s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
    for (int i = 0; i < per_thread; ++i) { volatile double* p = acc + (rng(s) >> 7) % span; *p = *p + 1.0; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = acc[0];
}
__global__ void atom_shared_u32(unsigned* out, int span, int per_thread) {  // native 32-bit shared atomics for scale
    extern __shared__ unsigned accu[];
    for (int i = threadIdx.x; i < span; i += blockDim.x) accu[i] = 0;
    __syncthreads();
    unsigned s = blockIdx.x * 9781u + threadIdx.x * 7919u + 17u;
    for (int i = 0; i < per_thread; ++i) atomicOr(accu + (rng(s) >> 7) % span, 1u << (i & 31));
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = accu[0];
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int i = 0; i < 3; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int per_thread = 4096;
    double* buf; cudaMalloc(&buf, (size_t)4 << 30); cudaMemset(buf, 0, (size_t)4 << 30);
    double* out; cudaMalloc(&out, 1 << 20);
    for (int bps : {2, 4, 8}) for (int threads : {256, 512}) {
        if (bps * threads > 2048) continue;
        int grid = sms * bps; double ops = (double)grid * threads * per_thread;
        for (size_t span : {(size_t)2500, (size_t)20000, (size_t)1 << 17, (size_t)1 << 20}) {
            if ((size_t)grid * span * 8 > ((size_t)4 << 30)) continue;
            float ms = timeit([&] { red_global<<<grid, threads>>>(buf, span, per_thread, 0); });
            printf("RED.F64 global  grid=%4d x%3d  slice=%8zu doubles (total %7.1f MB): %7.1f Gop/s\n", grid, threads, span, grid * span * 8 / 1e6, ops / ms / 1e6);
        }
        float ms = timeit([&] { red_global<<<grid, threads>>>(buf, (size_t)1 << 22, per_thread, 1); });
        printf("RED.F64 global  grid=%4d x%3d  all blocks share 32 MB: %7.1f Gop/s\n", grid, threads, ops / ms / 1e6);
    }
    for (int threads : {256, 512, 1024}) for (int span : {2048, 12288}) {
        int bps = 2048 / threads; size_t smem = (size_t)span * 8;
        if (bps * smem > 200000) bps = (int)(200000 / smem);
        int grid = sms * bps; double ops = (double)grid * threads * per_thread;
        cudaFuncSetAttribute(atom_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(plain_shared, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        float ms = timeit([&] { atom_shared<<<grid, threads, smem>>>(out, span, per_thread); });
        printf("ATOMS CAS f64   grid=%4d x%4d span=%6d: %7.1f Gop/s\n", grid, threads, span, ops / ms / 1e6);
        ms = timeit([&] { plain_shared<<<grid, threads, smem>>>(out, span, per_thread); });
        printf("plain smem RMW  grid=%4d x%4d span=%6d: %7.1f Gop/s\n", grid, threads, span, ops / ms / 1e6);
        ms = timeit([&] { atom_shared_u32<<<grid, threads, span * 4>>>((unsigned*)out, span, per_thread); });
        printf("ATOMS.OR u32    grid=%4d x%4d span=%6d: %7.1f Gop/s\n", grid, threads, span, ops / ms / 1e6);
    }
    return 0;
}
