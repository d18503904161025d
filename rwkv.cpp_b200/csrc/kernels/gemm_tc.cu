// Chunked-prefill contraction on the 5th-generation tensor cores:  Y[M, T] = epilogue(W[M, K] . X[K, T])  for T >= 32.
//
// This is the one place of the eval path where the chunk x embedding contraction is dense enough for tensor cores
// (~325 FLOP/B at T = 128). Per CTA: one 128-row tile of W against all T (<= 256) tokens.
//
//   A operand  the CTA's 256 threads read their own ggml block (or 32 f16 weights) of the tile from global memory
//              each K-step of 64, dequantise it to fp16 in registers ((q - offset) * d + m, rounded once) and store it
//              into shared memory in the canonical K-major no-swizzle UMMA layout (8-row x 16-byte core matrices)
//   B operand  fp16 activations [T][K] (converted once per launch by convert_f16_kernel), copied per K-step into the same
//              canonical layout
//   MMA        one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = T padded to 16, K = 16) four
//              times per K-step; the fp32 accumulator lives in TMEM (N columns x 128 lanes); tcgen05.commit signals an
//              mbarrier per shared-memory stage, so dequantisation of step k+1 overlaps the MMAs of step k
//   epilogue   warps 0-3 pull the accumulator with tcgen05.ld (32x32b), apply the same fused epilogues as the GEMV
//              (activation, bias, residual, gate) and store column-major.
//
// Numerics: weights exactly as the file stores them, activations rounded to fp16, fp32 accumulate -- what the reference
// does for F16 weights (ggml-cpu.c:259-264, 1463); for quantised weights the reference rounds activations to int8
// blocks instead, so this path is (slightly) more accurate than the reference, and NOT bit-identical to the decode GEMV.
// The engine therefore uses it only for passes of >= 32 tokens and never for F32 weights, which keeps every
// serial == sequence memcmp test of the reference (tests/test_eval_sequence_in_chunks.c: chunks of 1, 2, 8, 10) exact.
#include "gemv.h"
#include "quant_decode.cuh"

#include <cuda_fp16.h>
#include <cstring>

namespace rwkv {
namespace tc {

constexpr int TILE_M = 128;
constexpr int KSTEP = 64;                 // K elements per shared-memory stage = 4 MMAs of K = 16
constexpr int THREADS = 256;
constexpr int STAGES = 4;                // shared-memory ring; global loads run 2 K-steps ahead of the MMAs
constexpr int MAX_N = 256;

__device__ __forceinline__ uint32_t smem_u32(const void * p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t * bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// Bounded wait: a descriptor or TMEM mistake shows up as a trapped kernel, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t * bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; spins++) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (spins > (1u << 24)) __trap();
    }
}

// ---- tcgen05 wrappers (PTX ISA 8.6+, sm_100a) ----------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t * dst_in_smem, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t * bar) {   // arrives on `bar` when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t * r) {   // 32 lanes x 32 consecutive columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp:98-123;
// canonical layout ((8,n),2):((1,SBO),LBO) in 16-byte units, mma_traits_sm100.hpp:273-303):
//   element (row, k) lives at  start + (row % 8) * 16 + (row / 8) * SBO + (k / 8) * LBO   bytes
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t) ((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t) ((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t) ((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t) 1 << 46;     // descriptor version 1 (Blackwell)
    return d;                    // base_offset 0, lbo_mode 0, layout_type 0 = SWIZZLE_NONE
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor, mma_sm100_desc.hpp:412-439):
// D = f32, A = B = f16, both K-major, dense.
__host__ __device__ inline uint32_t make_idesc(int M, int N) {
    return (1u << 4) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}

// Tile bookkeeping: [A stage0][A stage1][B stage0][B stage1]
//   A stage: TILE_M x KSTEP halves = 16 KB : chunk (kc, g) at (kc * 16 + g) * 128 bytes   (kc = k / 8, g = row / 8)
//   B stage: NPAD   x KSTEP halves         : chunk (kc, g) at (kc * NG + g) * 128 bytes   (NG = NPAD / 8)
struct TcShared {
    uint64_t mma_done[STAGES];
    uint32_t tmem_base;
    GemvProblem P;
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }
__device__ __forceinline__ float tc_epilogue(const GemvProblem & P, int row, int col, float v) {
    switch (P.epi) {
        case EPI_SIGMOID: return sigmoidf_(v);
        case EPI_SILU: return v / (1.0f + expf(-v));
        case EPI_TANH: return tanhf(v);
        case EPI_RELU_SQR: { float r = fmaxf(v, 0.0f); return r * r; }
        case EPI_ADD: return P.res[(long long) col * P.ldres + row] + v;
        case EPI_MUL_ADD: return P.res[(long long) col * P.ldres + row] + P.gate[(long long) col * P.ldgate + row] * v;
        case EPI_BIAS_EXPNEGEXP: return expf(-expf(v + P.bias[row]));
        case EPI_BIAS_SIGMOID: return sigmoidf_(v + P.bias[row]);
        case EPI_BIAS_W7: return expf(sigmoidf_(v + P.bias[row]) * -0.606531f);
        default: return v;
    }
}

// 32 weights of one ggml block (or 32 f16 values) -> 4 x 16 bytes of fp16, element order 0..31
template <int TYPE> struct BlockRegs { uint32_t w[TYPE == DT_F16 ? 16 : (TYPE == DT_Q8_0 ? 9 : 6)]; };

template <int TYPE> __device__ __forceinline__ void load_block(const uint8_t * row, int blk, BlockRegs<TYPE> & r) {
    if constexpr (TYPE == DT_F16) {
        const uint4 * p = reinterpret_cast<const uint4 *>(row + (size_t) blk * 64);
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint4 v = __ldg(p + i); r.w[4 * i] = v.x; r.w[4 * i + 1] = v.y; r.w[4 * i + 2] = v.z; r.w[4 * i + 3] = v.w; }
    } else if constexpr (TYPE == DT_Q5_1) {          // 24-byte blocks on an 8-byte grid
        const uint2 * p = reinterpret_cast<const uint2 *>(row + (size_t) blk * 24);
#pragma unroll
        for (int i = 0; i < 3; i++) { const uint2 v = __ldg(p + i); r.w[2 * i] = v.x; r.w[2 * i + 1] = v.y; }
    } else if constexpr (TYPE == DT_Q4_1) {          // 20-byte blocks on a 4-byte grid
        const uint32_t * p = reinterpret_cast<const uint32_t *>(row + (size_t) blk * 20);
#pragma unroll
        for (int i = 0; i < 5; i++) r.w[i] = __ldg(p + i);
    } else {
        // 18 / 22 / 34-byte blocks start on a 2-byte grid: read whole words from the aligned-down address and realign
        constexpr int BB = QTraits<TYPE>::BLOCK_BYTES;
        constexpr int NW = BB / 4 + 1;
        const size_t start = (size_t) blk * BB;
        const uint32_t * p = reinterpret_cast<const uint32_t *>(row + (start & ~(size_t) 3));
        uint32_t t[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) t[i] = __ldg(p + i);
        if (start & 2) {
#pragma unroll
            for (int i = 0; i < NW - 1; i++) r.w[i] = funnel16(t[i], t[i + 1]);
        } else {
#pragma unroll
            for (int i = 0; i < NW - 1; i++) r.w[i] = t[i];
        }
        if (NW - 1 < (int) (sizeof(r.w) / 4)) r.w[NW - 1] = (start & 2) ? (t[NW - 1] >> 16) : t[NW - 1];
    }
}

// Dequantise one block to 32 fp16 values in element order with packed-half arithmetic: the stored integer goes into
// the mantissa of 1024.0h (0x6400 | q == 1024 + q exactly), one HSUB2 removes 1024 (+ the -8 / -16 offset of the
// symmetric formats, + 128 for the sign-flipped bytes of Q8_0), one HFMA2 applies w = q * d + m with a single rounding
// (dequantize_row_q*, ggml-quants.c:255-363, rounded to fp16). ~70 instructions per block instead of ~300 in fp32.
template <int TYPE> __device__ __forceinline__ void block_to_half(const BlockRegs<TYPE> & r, uint4 out[4]) {
    if constexpr (TYPE == DT_F16) {
#pragma unroll
        for (int i = 0; i < 4; i++) out[i] = make_uint4(r.w[4 * i], r.w[4 * i + 1], r.w[4 * i + 2], r.w[4 * i + 3]);
    } else {
        BlockQ bq;
        decode_block<TYPE>(r.w, 0, bq);          // q[0..3]: elements 0..15, q[4..7]: elements 16..31, 4 per word
        const __half2 d2 = __float2half2_rn(bq.d), m2 = __float2half2_rn(bq.m);
        const float bias = 1024.0f + (float) QTraits<TYPE>::OFFSET + (TYPE == DT_Q8_0 ? 128.0f : 0.0f);
        const __half2 bias2 = __float2half2_rn(bias);
        uint32_t * o = reinterpret_cast<uint32_t *>(out);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t q = (uint32_t) bq.q[i];
            if (TYPE == DT_Q8_0) q ^= 0x80808080u;                 // signed byte -> byte + 128
            const uint32_t p0 = __byte_perm(q, 0x64646464u, 0x5140);   // [b0, 0x64, b1, 0x64] = halves (1024 + b0, 1024 + b1)
            const uint32_t p1 = __byte_perm(q, 0x64646464u, 0x7362);   // [b2, 0x64, b3, 0x64]
            __half2 h0 = __hfma2(__hsub2(*reinterpret_cast<const __half2 *>(&p0), bias2), d2, m2);
            __half2 h1 = __hfma2(__hsub2(*reinterpret_cast<const __half2 *>(&p1), bias2), d2, m2);
            o[2 * i] = *reinterpret_cast<uint32_t *>(&h0);
            o[2 * i + 1] = *reinterpret_cast<uint32_t *>(&h1);
        }
    }
}

struct TcBatch {
    int n, T, npad;                  // problems, tokens, tokens padded to a multiple of 16
    int tmem_cols;                   // power of two >= 32
    const __half * act16[GEMV_MAX_PROBLEMS];   // [npad][K] fp16 per problem
    GemvProblem p[GEMV_MAX_PROBLEMS];          // first_cta / n_cta = tiles of 128 rows
    TraceRec * trace;
};

__device__ __forceinline__ void cp_async16(void * smem_dst, const void * gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// B tile of K-step `ks` (NPAD x 64 halves of fp16 activations) -> stage, asynchronously. 16 lanes cover 16 tokens, lane / 16
// picks one of two adjacent 8-element chunks, so a warp reads 32-byte runs and writes conflict-free 128-byte runs.
__device__ __forceinline__ void copy_b_async(uint8_t * b_stage, const __half * act16, int K, int NPAD, int k0) {
    const int NG = NPAD / 8;
    for (int i = threadIdx.x; i < NPAD * 8; i += THREADS) {
        const int pair = i / 32, lane = i % 32;
        const int tg = pair / 4, cp = pair % 4;
        const int n = tg * 16 + (lane % 16), kc = cp * 2 + lane / 16;
        cp_async16(b_stage + (uint32_t) (kc * NG + n / 8) * 128 + (n % 8) * 16, act16 + (size_t) n * K + k0 + kc * 8);
    }
}

template <int TYPE>
__device__ void tc_tile(TcShared & sh, uint8_t * smem, const TcBatch & batch, const __half * act16, int tile) {
    const GemvProblem & P = sh.P;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = P.K, NPAD = batch.npad, NG = NPAD / 8;
    const int row0 = tile * TILE_M;
    const uint32_t a_bytes = TILE_M * KSTEP * 2, b_bytes = (uint32_t) NPAD * KSTEP * 2;
    uint8_t * const a_base = smem;
    uint8_t * const b_base = smem + STAGES * a_bytes;

    // my weight block of every K-step: row (tid / 2), block (tid % 2) of the step
    const int my_row = min(row0 + tid / 2, P.M - 1);
    const uint8_t * wrow = reinterpret_cast<const uint8_t *>(P.W) + (size_t) my_row * (size_t) P.pitch;
    const int blk_in_step = tid & 1;
    const int r_local = tid / 2;
    const uint32_t a_dst0 = (uint32_t) ((blk_in_step * 4) * 16 + r_local / 8) * 128 + (r_local % 8) * 16;   // + c * 16 * 128 for chunk c

    const uint32_t idesc = make_idesc(TILE_M, NPAD);
    const int nsteps = K / KSTEP;
    // software pipeline, distance 2: weights of step ks+2 travel to registers and activations of step ks+2 to shared
    // memory (cp.async) while step ks is dequantised and multiplied
    // One K-step. `cur` holds this step's weight block, `fut` receives the block of step ks+2. The three register sets
    // rotate by NAME (the loop below is unrolled by 3): a `w0 = w1` style rotation would read the in-flight load at the end
    // of the very iteration that issued it and turn the distance-2 prefetch into distance 0.
    // phase accounting of CTA 0 / thread 0 (cycles), reported through the trace marks
    const bool acct = batch.trace != nullptr && blockIdx.x == 0 && tid == 0;
    long long acc_wait = 0, acc_deq = 0, acc_cp = 0, acc_sync = 0, tq = 0;
    auto tick = [&](long long & a) { if (acct) { const long long c = clock64(); a += c - tq; tq = c; } };
    auto step = [&](int ks, const BlockRegs<TYPE> & cur, BlockRegs<TYPE> & fut) {
        const int s = ks % STAGES;
        if (acct) tq = clock64();
        if (ks + 2 < nsteps) {
            const int s2 = (ks + 2) % STAGES;
            // stage s2 was last read by the MMAs of step ks-2
            if (ks >= 2) mbar_wait(&sh.mma_done[s2], (uint32_t) ((((ks - 2) / STAGES)) & 1));
            tick(acc_wait);
            load_block<TYPE>(wrow, (ks + 2) * 2 + blk_in_step, fut);
            copy_b_async(b_base + (size_t) s2 * b_bytes, act16, K, NPAD, (ks + 2) * KSTEP);
        }
        cp_async_commit();
        // A: dequantise my block of step ks -> 4 chunks of 8 halves. Stage s was last read at step ks-4, whose commit was
        // awaited at step ks-2.
        uint4 h[4];
        block_to_half<TYPE>(cur, h);
        uint8_t * a_stage = a_base + (size_t) s * a_bytes;
#pragma unroll
        for (int c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(a_stage + a_dst0 + (uint32_t) c * 16 * 128) = h[c];
        tick(acc_deq);
        cp_async_wait<2>();
        tick(acc_cp);          // the activations of step ks have landed (groups ks+1, ks+2 may still fly)
        fence_async_smem();          // generic-proxy writes -> visible to the tensor core's async proxy
        __syncthreads();
        tick(acc_sync);
        if (tid == 0) {
            tc_fence_after_sync();
            const uint32_t a_addr = smem_u32(a_stage), b_addr = smem_u32(b_base + (size_t) s * b_bytes);
#pragma unroll
            for (int j = 0; j < KSTEP / 16; j++) {
                const uint64_t adesc = make_desc(a_addr + (uint32_t) (2 * j) * 16 * 128, 16 * 128, 128);
                const uint64_t bdesc = make_desc(b_addr + (uint32_t) (2 * j) * NG * 128, (uint32_t) NG * 128, 128);
                umma_f16(sh.tmem_base, adesc, bdesc, idesc, (ks > 0 || j > 0) ? 1u : 0u);
            }
            umma_commit(&sh.mma_done[s]);
        }
    };
    BlockRegs<TYPE> w0, w1, w2;
    load_block<TYPE>(wrow, blk_in_step, w0);
    copy_b_async(b_base, act16, K, NPAD, 0);
    cp_async_commit();
    if (nsteps > 1) { load_block<TYPE>(wrow, 2 + blk_in_step, w1); copy_b_async(b_base + b_bytes, act16, K, NPAD, KSTEP); }
    cp_async_commit();
    for (int ks = 0; ks < nsteps; ks += 3) {
        step(ks, w0, w2);
        if (ks + 1 < nsteps) step(ks + 1, w1, w0);
        if (ks + 2 < nsteps) step(ks + 2, w2, w1);
    }
    if (acct && tile == (int) 0) {
        // marks = kernel start + cycles spent in: waiting for the MMA barrier | issue loads + dequantise (= waiting for the
        // weight block) | cp.async wait | __syncthreads
        TraceRec * t = batch.trace;
        t->mark[0] = t->start + (unsigned long long) acc_wait;
        t->mark[1] = t->start + (unsigned long long) acc_deq;
        t->mark[2] = t->start + (unsigned long long) acc_cp;
        t->mark[3] = t->start + (unsigned long long) acc_sync;
    }
    // the last commit covers every MMA of the tile
    {
        const int last = nsteps - 1;
        mbar_wait(&sh.mma_done[last % STAGES], (uint32_t) ((last / STAGES) & 1));
        tc_fence_after_sync();
    }
    // epilogue: warps 0..3 own TMEM lanes 32w .. 32w+31 = rows row0 + 32w + lane
    if (warp < 4) {
        const int row = row0 + warp * 32 + (tid & 31);
        for (int c0 = 0; c0 < NPAD; c0 += 32) {
            uint32_t acc[32];
            tmem_ld32(sh.tmem_base + ((uint32_t) (warp * 32) << 16) + (uint32_t) c0, acc);
            if (row < P.M) {
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const int col = c0 + j;
                    if (col < batch.T) P.y[(long long) col * P.ldy + row] = tc_epilogue(P, row, col, __uint_as_float(acc[j]));
                }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
}

__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const TcBatch batch) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ TcShared sh;
    trace_begin(batch.trace);
    int pi = 0;
    for (int i = 1; i < batch.n; i++) if ((int) blockIdx.x >= batch.p[i].first_cta) pi = i;
    if (threadIdx.x == 0) {
        sh.P = batch.p[pi];
        for (int s = 0; s < STAGES; s++) mbar_init(&sh.mma_done[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(&sh.tmem_base, (uint32_t) batch.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    pdl_prologue();     // activations (act16, residuals) come from the previous kernels
    const int tile = (int) blockIdx.x - sh.P.first_cta;
    const __half * act16 = batch.act16[pi];
    switch (sh.P.type) {
        case DT_Q4_0: tc_tile<DT_Q4_0>(sh, smem, batch, act16, tile); break;
        case DT_Q4_1: tc_tile<DT_Q4_1>(sh, smem, batch, act16, tile); break;
        case DT_Q5_0: tc_tile<DT_Q5_0>(sh, smem, batch, act16, tile); break;
        case DT_Q5_1: tc_tile<DT_Q5_1>(sh, smem, batch, act16, tile); break;
        case DT_Q8_0: tc_tile<DT_Q8_0>(sh, smem, batch, act16, tile); break;
        default: tc_tile<DT_F16>(sh, smem, batch, act16, tile); break;
    }
    if (threadIdx.x < 32) tmem_dealloc(sh.tmem_base, (uint32_t) batch.tmem_cols);
    trace_end(batch.trace);
}

// x fp32 [K, T] column-major (column t contiguous) -> fp16 [npad][K], rows >= T zero-filled
__global__ void convert_f16_kernel(const float * x, long long ldx, int K, int T, int npad, __half * out, TraceRec * trace) {
    trace_begin(trace);
    pdl_prologue();
    const long long n4 = (long long) npad * K / 4;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long) gridDim.x * blockDim.x) {
        const long long e = i * 4;
        const int t = (int) (e / K), k = (int) (e % K);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t < T) v = *reinterpret_cast<const float4 *>(x + (long long) t * ldx + k);
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        *reinterpret_cast<uint2 *>(out + e) = make_uint2(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b));
    }
    trace_end(trace);
}

}  // namespace tc

bool gemm_tc_supported(const GemvProblem & p, int T) {
    return T >= 32 && T <= tc::MAX_N && p.type != DT_F32 && p.K % tc::KSTEP == 0 && p.K >= tc::KSTEP && (p.ldx % 4) == 0 &&
           (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
}

// act16_scratch: device buffer of at least sum over problems of npad * K halves.
cudaError_t gemm_tc_launch(GemvBatch & batch, const DeviceInfo & dev, cudaStream_t stream, void * act16_scratch, size_t scratch_bytes) {
    (void) dev;
    tc::TcBatch tb;
    memset(&tb, 0, sizeof(tb));
    tb.n = batch.n; tb.T = batch.T;
    tb.npad = (batch.T + 15) / 16 * 16;
    tb.tmem_cols = 32;
    while (tb.tmem_cols < tb.npad) tb.tmem_cols *= 2;
    __half * scratch = reinterpret_cast<__half *>(act16_scratch);
    size_t used = 0;
    int next = 0;
    for (int i = 0; i < batch.n; i++) {
        GemvProblem & p = batch.p[i];
        const size_t need = (size_t) tb.npad * p.K * sizeof(__half);
        // problems that share an input share the converted copy
        const __half * shared = nullptr;
        for (int j = 0; j < i; j++) if (batch.p[j].x == p.x && batch.p[j].K == p.K && batch.p[j].ldx == p.ldx) shared = tb.act16[j];
        if (!shared) {
            if (used + need > scratch_bytes) return cudaErrorMemoryAllocation;
            __half * dst = scratch + used / sizeof(__half);
            const long long n4 = (long long) tb.npad * p.K / 4;
            const int blocks = (int) ((n4 + 255) / 256 < 1184 ? (n4 + 255) / 256 : 1184);
            g_kernel_launches++;
            cudaError_t e = launch_pdl(tc::convert_f16_kernel, dim3(blocks), dim3(256), 0, stream, p.x, p.ldx, p.K, batch.T, tb.npad, dst, trace_slot("convert_f16"));
            if (e != cudaSuccess) return e;
            shared = dst;
            used += (need + 255) & ~(size_t) 255;
        }
        tb.act16[i] = shared;
        p.first_cta = next;
        p.n_cta = (p.M + tc::TILE_M - 1) / tc::TILE_M;
        next += p.n_cta;
        tb.p[i] = p;
    }
    const size_t smem = (size_t) tc::STAGES * tc::TILE_M * tc::KSTEP * 2 + (size_t) tc::STAGES * tb.npad * tc::KSTEP * 2;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc::gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::STAGES * tc::TILE_M * tc::KSTEP * 2 + tc::STAGES * tc::MAX_N * tc::KSTEP * 2);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    tb.trace = trace_slot("gemm_tc");
    g_kernel_launches++;
    return launch_pdl(tc::gemm_tc_kernel, dim3(next), dim3(tc::THREADS), smem, stream, tb);
}

}  // namespace rwkv
