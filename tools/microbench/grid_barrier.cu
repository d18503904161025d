// Micro-benchmark: latency of a grid-wide barrier between the 2 x SMs CTAs of a persistent kernel on this GPU, for the variants
// the persistent decode kernel (rwkv.cpp_b200/csrc/kernels/decode_persistent.cu) could use. One number per variant: microseconds
// per barrier, measured with CUDA events around a cooperative launch that does nothing but N barriers (plus, optionally, one
// dependent global load after each barrier, which is what a real phase starts with).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/grid_barrier tools/microbench/grid_barrier.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned long long ld_acquire(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed(const unsigned long long * p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release(unsigned long long * p) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

// variant 0: fence + atomicAdd, ld.acquire poll, fence (what decode_persistent.cu does)
// variant 1: red.release, ld.acquire poll (no explicit fences)
// variant 2: fence + atomicAdd, ld.relaxed poll, fence
// variant 3: cooperative_groups grid.sync()
// variant 4: variant 1 with arrival counters spread over 8 cache lines (CTA b adds to line b % 8; the poller sums them)
template <int VARIANT, bool TOUCH>
__global__ void __launch_bounds__(288, 2) barrier_kernel(unsigned long long * counters, int n, float * data, float * sink) {
    cg::grid_group grid = cg::this_grid();
    const int tid = threadIdx.x;
    float acc = 0.f;
    for (int i = 1; i <= n; i++) {
        if (TOUCH) data[(size_t) blockIdx.x * 288 + tid] = acc + (float) i;      // a store the barrier has to publish
        if (VARIANT == 3) {
            grid.sync();
        } else {
            __syncthreads();
            if (tid == 0) {
                const unsigned long long target = (unsigned long long) i * gridDim.x;
                if (VARIANT == 0 || VARIANT == 2) { __threadfence(); atomicAdd(counters, 1ull); }
                else if (VARIANT == 1) red_release(counters);
                else red_release(counters + 16 * (blockIdx.x & 7));
                if (VARIANT == 0 || VARIANT == 1) { while (ld_acquire(counters) < target) {} }
                else if (VARIANT == 2) { while (ld_relaxed(counters) < target) {} }
                else {
                    for (;;) {
                        unsigned long long s = 0;
#pragma unroll
                        for (int k = 0; k < 8; k++) s += ld_acquire(counters + 16 * k);
                        if (s >= target) break;
                    }
                }
                if (VARIANT == 0 || VARIANT == 2) __threadfence();
            }
            __syncthreads();
        }
        if (TOUCH) acc += data[(size_t) ((blockIdx.x + 1) % gridDim.x) * 288 + tid];   // first dependent load of the next phase
    }
    if (TOUCH && acc == -1.f) *sink = acc;
}

template <int VARIANT, bool TOUCH>
static float run(int grid, int n, unsigned long long * counters, float * data, float * sink) {
    cudaMemset(counters, 0, 4096);
    void * args[] = {&counters, &n, &data, &sink};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaLaunchCooperativeKernel((const void *) barrier_kernel<VARIANT, TOUCH>, dim3(grid), dim3(288), args, 0, 0);   // warm-up
    cudaDeviceSynchronize();
    cudaMemset(counters, 0, 4096);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchCooperativeKernel((const void *) barrier_kernel<VARIANT, TOUCH>, dim3(grid), dim3(288), args, 0, 0);
    cudaEventRecord(e1);
    if (e != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) { printf("variant %d failed: %s\n", VARIANT, cudaGetErrorString(cudaGetLastError())); return -1.f; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms * 1e3f / (float) n;
}

int main(int argc, char ** argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 2000;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("no CUDA device\n"); return 1; }
    const int grid = 2 * prop.multiProcessorCount;
    unsigned long long * counters; float * data, * sink;
    cudaMalloc(&counters, 4096); cudaMalloc(&data, (size_t) grid * 288 * 4); cudaMalloc(&sink, 4);
    cudaMemset(data, 0, (size_t) grid * 288 * 4);
    printf("%s, %d CTAs x 288 threads, %d barriers per launch; microseconds per barrier\n", prop.name, grid, n);
    printf("variant                                   bare    +store/load\n");
    printf("0 fence+atomicAdd / ld.acquire / fence   %6.3f   %6.3f\n", run<0, false>(grid, n, counters, data, sink), run<0, true>(grid, n, counters, data, sink));
    printf("1 red.release / ld.acquire               %6.3f   %6.3f\n", run<1, false>(grid, n, counters, data, sink), run<1, true>(grid, n, counters, data, sink));
    printf("2 fence+atomicAdd / ld.relaxed / fence   %6.3f   %6.3f\n", run<2, false>(grid, n, counters, data, sink), run<2, true>(grid, n, counters, data, sink));
    printf("3 cooperative_groups grid.sync()         %6.3f   %6.3f\n", run<3, false>(grid, n, counters, data, sink), run<3, true>(grid, n, counters, data, sink));
    printf("4 red.release on 8 lines / 8 x ld.acquire %6.3f   %6.3f\n", run<4, false>(grid, n, counters, data, sink), run<4, true>(grid, n, counters, data, sink));
    return 0;
}
