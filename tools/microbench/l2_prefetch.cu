// Micro-benchmark: can the 126 MB L2 serve as the landing zone of a weight stream that runs AHEAD of the kernels that consume it?
// The decode step is a chain of ~9 dependent launches per layer; each leaves HBM idle for the 4-7 us of its dependency bubble
// (profiles/r2_trace_decode_*.csv). If the NEXT launch's weights can be pulled into L2 during that bubble (cp.async.bulk.prefetch.L2,
// no registers, no shared memory, fire and forget), the consuming launch runs at L2 speed and HBM never idles. This measures, on
// this GPU:
//   1. streaming-read bandwidth of X MB cold (HBM) vs after a completed L2 prefetch of the same X MB, X = 8..160 MB
//      -> how much prefetched data survives, and the L2 -> SM rate
//   2. the same when the prefetch of buffer B is issued while another kernel streams buffer A from HBM (interference, both ways)
//   3. how long a prefetch of X MB takes to land (poll by timing a read issued d microseconds after the prefetch)
// Reads use the same mechanism as the product kernel: 1-D bulk copies (cp.async.bulk.shared::cluster.global) through a 3-stage
// mbarrier ring, 2 CTAs per SM, L2::evict_first on the stream.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/l2_prefetch tools/microbench/l2_prefetch.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t * bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ uint64_t policy_evict_last() { uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p; }
__device__ __forceinline__ void bulk_g2s(void * dst, const void * src, uint32_t bytes, uint64_t * bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void * src, uint32_t bytes) { asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory"); }
__device__ __forceinline__ void prefetch_l2_hint(const void * src, uint32_t bytes, uint64_t policy) {
    asm volatile("cp.async.bulk.prefetch.L2.global.L2::cache_hint [%0], %1, %2;" ::"l"(src), "r"(bytes), "l"(policy) : "memory");
}

constexpr int STAGE = 30 * 1024, NST = 3, THREADS = 288;

// Streams [src, src + bytes) through shared memory in STAGE-byte tiles, tile t to CTA t % gridDim.x (the product kernel's pattern);
// the 8 consumer warps read every word (LDS.128) so the data really has to arrive.
__global__ void __launch_bounds__(THREADS, 2) stream_kernel(const uint8_t * src, size_t bytes, int evict_first, unsigned * sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t full[NST], empty[NST];
    const size_t ntiles = (bytes + STAGE - 1) / STAGE;
    const int mine = (int) ((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    if (threadIdx.x == 0) { for (int s = 0; s < NST; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 256) {
        const uint64_t pol = evict_first ? policy_evict_first() : policy_evict_last();
        for (int i = 0; i < mine; i++) {
            const int s = i % NST;
            if (i >= NST) mbar_wait(&empty[s], (uint32_t) (((i / NST) - 1) & 1));
            const size_t off = ((size_t) blockIdx.x + (size_t) i * gridDim.x) * STAGE;
            const uint32_t n = (uint32_t) (bytes - off < STAGE ? bytes - off : STAGE);
            mbar_expect_tx(&full[s], n);
            bulk_g2s(smem + s * STAGE, src + off, n, &full[s], pol);
        }
    } else if (threadIdx.x < 256) {
        unsigned acc = 0;
        for (int i = 0; i < mine; i++) {
            const int s = i % NST;
            mbar_wait(&full[s], (uint32_t) ((i / NST) & 1));
            const uint4 * p = reinterpret_cast<const uint4 *>(smem + s * STAGE);
            for (int k = threadIdx.x; k < STAGE / 16; k += 256) { const uint4 v = p[k]; acc += v.x ^ v.y ^ v.z ^ v.w; }
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty[s]);
        }
        if (acc == 0x12345678u) sink[0] = acc;
    }
}

// One thread per 16 KB: fire-and-forget L2 prefetch of [src, src + bytes). hint: 0 none, 1 evict_last, 2 evict_first
__global__ void prefetch_kernel(const uint8_t * src, size_t bytes, int hint) {
    const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
    const size_t off = i * 16384;
    if (off >= bytes) return;
    const uint32_t n = (uint32_t) (bytes - off < 16384 ? bytes - off : 16384);
    if (hint == 0) prefetch_l2(src + off, n);
    else prefetch_l2_hint(src + off, n, hint == 1 ? policy_evict_last() : policy_evict_first());
}
__global__ void spin_kernel(long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while ((long long) (t - t0) < ns);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount, grid = 2 * sms;
    printf("%s, %d SMs, L2 %.0f MB\n", prop.name, sms, prop.l2CacheSize / 1e6);
    const size_t POOL = (size_t) 3 << 30;      // rotate through 3 GB so nothing is warm by accident
    uint8_t * pool; unsigned * sink;
    CK(cudaMalloc(&pool, POOL)); CK(cudaMemset(pool, 1, POOL)); CK(cudaMalloc(&sink, 64));
    CK(cudaFuncSetAttribute(stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NST * STAGE));
    cudaStream_t s1, s2; CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    cudaEvent_t e0, e1, e2, e3; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2)); CK(cudaEventCreate(&e3));
    size_t cursor = 0;
    auto fresh = [&](size_t bytes) { if (cursor + bytes > POOL) cursor = 0; uint8_t * p = pool + cursor; cursor += (bytes + (1 << 21) - 1) & ~(size_t) ((1 << 21) - 1); return p; };
    auto stream_ms = [&](const uint8_t * p, size_t bytes, cudaStream_t st) {
        CK(cudaEventRecord(e0, st));
        stream_kernel<<<grid, THREADS, NST * STAGE, st>>>(p, bytes, 1, sink);
        CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); return ms;
    };
    for (int i = 0; i < 3; i++) stream_ms(fresh(256 << 20), 256 << 20, s1);       // warm-up
    printf("cold stream 256 MB: %.0f GB/s\n", (256 << 20) / 1e6 / stream_ms(fresh(256 << 20), 256 << 20, s1));

    printf("\n== 1. prefetch X MB, let it land (200 us), then stream it: GB/s (cold reference beside it)\n");
    const int sizes[] = {8, 16, 32, 48, 64, 80, 96, 112, 128, 160};
    for (int hint = 0; hint < 2; hint++) {
        for (int mb : sizes) {
            const size_t bytes = (size_t) mb << 20;
            float best_cold = 1e9f, best_hit = 1e9f;
            for (int rep = 0; rep < 3; rep++) {
                best_cold = fminf(best_cold, stream_ms(fresh(bytes), bytes, s1));
                uint8_t * p = fresh(bytes);
                prefetch_kernel<<<(unsigned) ((bytes / 16384 + 255) / 256), 256, 0, s1>>>(p, bytes, hint);
                spin_kernel<<<1, 1, 0, s1>>>(200000);
                best_hit = fminf(best_hit, stream_ms(p, bytes, s1));
            }
            printf("  hint %d  %4d MB: cold %6.0f GB/s (%6.1f us)   after prefetch %6.0f GB/s (%6.1f us)\n", hint, mb, bytes / 1e6 / best_cold, best_cold * 1e3, bytes / 1e6 / best_hit, best_hit * 1e3);
        }
    }

    printf("\n== 2. how fast does a prefetch land? prefetch X MB, wait d us, stream it (us for the stream; cold = no prefetch)\n");
    for (int mb : {16, 48}) {
        const size_t bytes = (size_t) mb << 20;
        printf("  %d MB: cold %.1f us;", mb, stream_ms(fresh(bytes), bytes, s1) * 1e3);
        for (int d : {0, 2, 4, 8, 16, 32}) {
            float best = 1e9f;
            for (int rep = 0; rep < 3; rep++) {
                uint8_t * p = fresh(bytes);
                prefetch_kernel<<<(unsigned) ((bytes / 16384 + 255) / 256), 256, 0, s1>>>(p, bytes, 0);
                if (d) spin_kernel<<<1, 1, 0, s1>>>(d * 1000ll);
                best = fminf(best, stream_ms(p, bytes, s1));
            }
            printf("  d=%d: %.1f", d, best * 1e3);
        }
        printf("\n");
    }

    printf("\n== 3. stream A (cold, HBM) while B is being prefetched on another stream; then stream B\n");
    for (int mb : {16, 48}) {
        const size_t bytes = (size_t) mb << 20;
        for (int rep = 0; rep < 2; rep++) {
            uint8_t * a = fresh(bytes), * b = fresh(bytes);
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s1));
            prefetch_kernel<<<(unsigned) ((bytes / 16384 + 255) / 256), 256, 0, s2>>>(b, bytes, 0);
            stream_kernel<<<grid, THREADS, NST * STAGE, s1>>>(a, bytes, 1, sink);
            CK(cudaEventRecord(e1, s1));
            stream_kernel<<<grid, THREADS, NST * STAGE, s1>>>(b, bytes, 1, sink);
            CK(cudaEventRecord(e2, s1)); CK(cudaEventSynchronize(e2));
            float ta, tb; CK(cudaEventElapsedTime(&ta, e0, e1)); CK(cudaEventElapsedTime(&tb, e1, e2));
            printf("  %d MB: A (with B's prefetch in flight) %.1f us = %.0f GB/s;  B afterwards %.1f us = %.0f GB/s;  A+B %.1f us vs 2 x cold\n", mb, ta * 1e3, bytes / 1e6 / ta, tb * 1e3,
                   bytes / 1e6 / tb, (ta + tb) * 1e3);
        }
    }

    printf("\n== 4. chain of 8 'launches' of X MB each: plain vs each launch preceded by the prefetch of the NEXT one (same stream, prefetch kernel first)\n");
    for (int mb : {12, 48}) {
        const size_t bytes = (size_t) mb << 20;
        for (int mode = 0; mode < 2; mode++) {
            std::vector<uint8_t *> bufs; for (int i = 0; i < 9; i++) bufs.push_back(fresh(bytes));
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s1));
            for (int i = 0; i < 8; i++) {
                if (mode) prefetch_kernel<<<(unsigned) ((bytes / 16384 + 255) / 256), 256, 0, s1>>>(bufs[i + 1], bytes, 0);
                stream_kernel<<<grid, THREADS, NST * STAGE, s1>>>(bufs[i], bytes, 1, sink);
                spin_kernel<<<1, 1, 0, s1>>>(5000);                    // the dependency bubble of the real chain
            }
            CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("  %d MB x 8, %s: %.1f us per launch (incl. 5 us bubble)\n", mb, mode ? "next launch prefetched" : "plain", ms * 1e3 / 8);
        }
    }
    return 0;
}
