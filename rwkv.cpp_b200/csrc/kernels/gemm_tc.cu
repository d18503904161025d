// Chunked-prefill contraction on the 5th-generation tensor cores:  Y[M, T] = epilogue(W[M, K] . X[K, T])  for T >= 32.
//
// This is the one place of the eval path where the chunk x embedding contraction is dense enough for tensor cores
// (~325 FLOP/B at T = 128). Per CTA: one 128-row tile of W against all T (<= 256) tokens, warp-specialised:
//
//   warp 8   raw producer   one contiguous bulk copy (cp.async.bulk, 50 KB) per K-chunk (8 K-steps: 128 rows x 16 ggml blocks,
//                           still quantised) out of the TILE-MAJOR prefill copy of the matrix (tc_repack_kernel) into a
//                           2-deep shared-memory ring; weights do not depend on the previous kernel, so this starts
//                           before the programmatic-dependency wait. (A 2-D TMA box over the row-major matrix needed
//                           ~58 cycles per 416-byte row: 3.7 us per chunk, which capped the K loop at ~1000 cycles per step.)
//   warp 9   B producer     fp16 activations, stored by convert_f16_kernel directly in the UMMA canonical layout, one
//                           contiguous bulk copy (16-32 KB) per K-step of 64
//   warps 0-7 transform     each thread owns (row, block) of the K-step: read the raw block from shared memory, dequantise
//                           to fp16 in packed-half arithmetic ((q - offset) * d + m, rounded once), store into the A stage
//                           in the canonical K-major no-swizzle layout (8-row x 16-byte core matrices), proxy fence,
//                           arrive. These threads never have a global load in flight, so the fence costs nothing (it is a
//                           MEMBAR.ALL.CTA in SASS: with in-flight LDGs it serialised every K-step at HBM latency)
//   warp 10  MMA issuer     one thread issues tcgen05.mma.cta_group::1.kind::f16 (M = 128, N = T padded to 16, K = 16) four
//                           times per K-step; the fp32 accumulator lives in TMEM (N columns x 128 lanes); tcgen05.commit
//                           frees the A/B stage through an mbarrier
//   epilogue warps 0-3 pull the accumulator with tcgen05.ld (32x32b), apply the same fused epilogues as the GEMV
//                           (activation, bias, residual, gate) and store column-major.
//
// There is no __syncthreads in the K loop: every hand-off is an mbarrier (raw_full/raw_empty, a_full/b_full/ab_empty).
//
// Numerics: weights exactly as the file stores them, activations rounded to fp16, fp32 accumulate -- what the reference
// does for F16 weights (ggml-cpu.c:259-264, 1463); for quantised weights the reference rounds activations to int8
// blocks instead, so this path is (slightly) more accurate than the reference, and NOT bit-identical to the decode GEMV.
// The engine therefore uses it only for passes of >= 32 tokens and never for F32 weights, which keeps every
// serial == sequence memcmp test of the reference (tests/test_eval_sequence_in_chunks.c: chunks of 1, 2, 8, 10) exact.
#include "gemv.h"
#include "quant_decode.cuh"

#include <cuda_fp16.h>
#include <cstdlib>
#include <cstring>

namespace rwkv {
namespace tc {

constexpr int TILE_M = 128;
constexpr int KSTEP = 64;                 // K elements per A/B stage = 4 MMAs of K = 16
constexpr int XFORM_WARPS = 8;
constexpr int WARP_RAW = 8, WARP_B = 9, WARP_MMA = 10;
constexpr int THREADS = 352;
constexpr int MAX_STAGES = 8;             // A/B ring depth (runtime: as many B stages as fit shared memory, <= 8). A stage = 32 columns of
                                          // TENSOR MEMORY (dequantised weights) + NPAD x 64 halves of shared memory (activations)
constexpr int MAX_RAW_STAGES = 3;         // raw (quantised) ring, in K-chunks
constexpr int MAX_N = 256;
// Tile-major prefill copy of a matrix: [tile of 128 rows][K-chunk][row][ROW_STRIDE bytes]; ROW_STRIDE = chunk bytes + a
// pad that puts the 32 rows a warp reads on distinct shared-memory banks (64-bit loads for Q5_1, 128-bit for F16, 32-bit
// for the rest), so one chunk of one tile is ONE contiguous, bank-conflict-free bulk copy.
constexpr int QUANT_CHUNK_STEPS = 4;       // K-steps per raw chunk of the block formats (18-35 KB per bulk copy)
template <int TYPE> struct RawTraits {
    static constexpr int BLOCK_BYTES = QTraits<TYPE>::BLOCK_BYTES;
    static constexpr int CHUNK_STEPS = QUANT_CHUNK_STEPS;
    static constexpr int CHUNK_BYTES = CHUNK_STEPS * 2 * BLOCK_BYTES;
    static constexpr int ROW_STRIDE = CHUNK_BYTES + (TYPE == DT_Q5_1 ? 8 : 4);
};
template <> struct RawTraits<DT_F16> {
    static constexpr int BLOCK_BYTES = 64, CHUNK_STEPS = 2, CHUNK_BYTES = 256, ROW_STRIDE = CHUNK_BYTES + 16;
};
struct RawGeom { int block_bytes, chunk_steps, chunk_bytes, row_stride; };
inline RawGeom raw_geom(int type) {
    RawGeom g;
    g.block_bytes = type == DT_F16 ? 64 : dtype_block_bytes(type);
    g.chunk_steps = type == DT_F16 ? 2 : QUANT_CHUNK_STEPS;
    g.chunk_bytes = g.chunk_steps * 2 * g.block_bytes;
    g.row_stride = g.chunk_bytes + (type == DT_F16 ? 16 : type == DT_Q5_1 ? 8 : 4);
    return g;
}

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
        if (spins > 256) __nanosleep(32);          // a waiting role must not steal issue slots from a co-resident kernel
        if (spins > (1u << 24)) __trap();
    }
}

__device__ __forceinline__ void mbar_arrive(uint64_t * bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t * bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void * dst, const void * src, uint32_t bytes, uint64_t * bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
// ---- thread-block clusters: the CTAs of a cluster work on neighbouring row tiles of the SAME matrix over the same K range, so they
// need the same B (activation) stage; each loads 1/CS of it and multicasts the slice into every member's shared memory.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void bulk_load_multicast(void * dst, const void * src, uint32_t bytes, uint64_t * bar, uint16_t mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t * bar, uint32_t cta) {      // the barrier at the same offset in CTA `cta` of the cluster
    asm volatile(
        "{\n"
        ".reg .b32 ra;\n"
        "mapa.shared::cluster.u32 ra, %0, %1;\n"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint2 lds64(uint32_t a) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a)); return v; }
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a)); return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
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
// A operand from tensor memory (128 lanes x 8 columns of packed fp16 per K = 16), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes (this warp's quarter of TMEM) x 32 consecutive columns <- registers; r[j] goes to column j of the lane's row
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t * r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t * bar) {   // arrives on `bar` when every MMA issued so far has completed
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_commit_multicast(uint64_t * bar, uint16_t mask) {   // ... on the barrier at this offset in every CTA of `mask`
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
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

// Dynamic shared memory: [B stages][raw stages]
//   B stage: NPAD x KSTEP halves : chunk (kc, g) at (kc * NG + g) * 128 bytes   (kc = k / 8, g = token / 8)
//   raw stage: TILE_M rows x ROW_STRIDE bytes of quantised blocks, exactly as the tile-major copy stores them
// Tensor memory: columns [0, NPAD) fp32 accumulator, then up to MAX_STAGES x 32 columns of A: lane = tile row, column j of a stage =
//   fp16 pair (k = 2j, 2j + 1) of the K-step. The tensor core reads A from there (tcgen05.mma with a TMEM A operand), so
//   the dequantised weights never touch shared memory: its bandwidth (128 B/clk) is left to the B operand and the TMA
//   writes -- with A in shared memory the MMAs, the transform stores and the TMA traffic together needed ~600 cycles of
//   shared-memory time per K-step against 262 cycles of tensor-core math.
struct TcShared {
    uint64_t raw_full[MAX_RAW_STAGES], raw_empty[MAX_RAW_STAGES];
    uint64_t a_full[MAX_STAGES], b_full[MAX_STAGES], ab_empty[MAX_STAGES];
    uint64_t acc_done;
    uint32_t tmem_base;
    long long t0;            // clock64 at kernel start (trace marks)
    GemvProblem P;
    float colscale[MAX_N];
    int split_rank;          // arrival order of this CTA among the K-splits of its tile (split-K epilogue)
};

__device__ __forceinline__ float sigmoidf_(float v) { return 1.0f / (1.0f + expf(-v)); }

// 32 accumulator columns of one row -> fused epilogue -> column-major store. Everything the loop needs sits in registers
// (the problem record lives in shared memory: re-reading it per element made the unrolled body 250 instructions long).
struct EpiRow { float * y; long long ldy; const float * res; long long ldres; const float * gate; long long ldgate; float bias; };
template <int EPI> __device__ __forceinline__ void store_cols(const EpiRow & e, const uint32_t (&acc)[32], const float * cs, int c0, int T) {
    // the residual / gate inputs of all 32 columns are requested before the first store: the output may alias the residual
    // (x += ...), so a load placed after a store would have to wait for it, one L2 round trip per column
    float rv[32], gv[32];
    if constexpr (EPI == EPI_ADD || EPI == EPI_MUL_ADD) {
#pragma unroll
        for (int j = 0; j < 32; j++) rv[j] = c0 + j < T ? e.res[(long long) (c0 + j) * e.ldres] : 0.f;
    }
    if constexpr (EPI == EPI_MUL_ADD) {
#pragma unroll
        for (int j = 0; j < 32; j++) gv[j] = c0 + j < T ? e.gate[(long long) (c0 + j) * e.ldgate] : 0.f;
    }
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const int col = c0 + j;
        if (col >= T) break;
        float v = __uint_as_float(acc[j]) * cs[col];
        if constexpr (EPI == EPI_SIGMOID) v = sigmoidf_(v);
        else if constexpr (EPI == EPI_SILU) v = v / (1.0f + expf(-v));
        else if constexpr (EPI == EPI_TANH) v = tanhf(v);
        else if constexpr (EPI == EPI_RELU_SQR) { const float r = fmaxf(v, 0.0f); v = r * r; }
        else if constexpr (EPI == EPI_ADD) v = rv[j] + v;
        else if constexpr (EPI == EPI_MUL_ADD) v = rv[j] + gv[j] * v;
        else if constexpr (EPI == EPI_BIAS_EXPNEGEXP) v = expf(-expf(v + e.bias));
        else if constexpr (EPI == EPI_BIAS_SIGMOID) v = sigmoidf_(v + e.bias);
        else if constexpr (EPI == EPI_BIAS_W7) v = expf(sigmoidf_(v + e.bias) * -0.606531f);
        e.y[(long long) col * e.ldy] = v;
    }
}
// warp w owns TMEM lanes 32 (w % 4) .. + 31 = rows row0 + 32 (w % 4) + lane; warps 0-3 take the even 32-column groups of the
// accumulator, warps 4-7 the odd ones.
//   MODE 0  single split: accumulator -> fused epilogue -> y
//   MODE 1  one of several K-splits: accumulator -> this split's slot of the partial buffer (row-major [128][npad rounded up to 32], full-line stores)
//   MODE 2  the split that arrived last: partials of ALL splits, added in split order from the buffer -> fused epilogue -> y
template <int MODE>
__device__ __noinline__ void tc_epilogue_rows(const GemvProblem & Psh, const float * cs, uint32_t tmem_base, int row0, int npad, int T,
                                              float * part0, int nsplit, size_t slot_floats) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q = warp & 3;
    const int row = row0 + q * 32 + lane;
    const bool live = row < Psh.M;
    const int epi = Psh.epi;
    EpiRow e;
    e.y = Psh.y + row; e.ldy = Psh.ldy;
    e.res = Psh.res ? Psh.res + row : nullptr; e.ldres = Psh.ldres;
    e.gate = Psh.gate ? Psh.gate + row : nullptr; e.ldgate = Psh.ldgate;
    e.bias = (MODE != 1 && live && Psh.bias) ? Psh.bias[row] : 0.f;
    // partial rows are (npad rounded up to 32) floats apart: the loop below moves whole 32-column groups, and with a stride of npad the
    // last group of a row would spill into the next row's partial whenever npad is not a multiple of 32
    float * prow = part0 + (size_t) (q * 32 + lane) * (size_t) ((npad + 31) & ~31);
#pragma unroll 1
    for (int c0 = (warp >> 2) * 32; c0 < npad; c0 += 64) {
        uint32_t acc[32];
        if constexpr (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 32; j++) acc[j] = 0u;          // +0.0f
            for (int sp = 0; sp < nsplit; sp++) {
                const float4 * src = reinterpret_cast<const float4 *>(prow + (size_t) sp * slot_floats + c0);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float4 v = __ldcg(src + j);            // written by other SMs: L2, not L1
                    acc[4 * j] = __float_as_uint(__uint_as_float(acc[4 * j]) + v.x); acc[4 * j + 1] = __float_as_uint(__uint_as_float(acc[4 * j + 1]) + v.y);
                    acc[4 * j + 2] = __float_as_uint(__uint_as_float(acc[4 * j + 2]) + v.z); acc[4 * j + 3] = __float_as_uint(__uint_as_float(acc[4 * j + 3]) + v.w);
                }
            }
        } else {
            tmem_ld32(tmem_base + ((uint32_t) (q * 32) << 16) + (uint32_t) c0, acc);
        }
        if constexpr (MODE == 1) {
            float4 * dst = reinterpret_cast<float4 *>(prow + c0);
#pragma unroll
            for (int j = 0; j < 8; j++)
                __stcg(dst + j, make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]), __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3])));
            continue;
        }
        if (!live) continue;
        switch (epi) {
            case EPI_SIGMOID: store_cols<EPI_SIGMOID>(e, acc, cs, c0, T); break;
            case EPI_SILU: store_cols<EPI_SILU>(e, acc, cs, c0, T); break;
            case EPI_TANH: store_cols<EPI_TANH>(e, acc, cs, c0, T); break;
            case EPI_RELU_SQR: store_cols<EPI_RELU_SQR>(e, acc, cs, c0, T); break;
            case EPI_ADD: store_cols<EPI_ADD>(e, acc, cs, c0, T); break;
            case EPI_MUL_ADD: store_cols<EPI_MUL_ADD>(e, acc, cs, c0, T); break;
            case EPI_BIAS_EXPNEGEXP: store_cols<EPI_BIAS_EXPNEGEXP>(e, acc, cs, c0, T); break;
            case EPI_BIAS_SIGMOID: store_cols<EPI_BIAS_SIGMOID>(e, acc, cs, c0, T); break;
            case EPI_BIAS_W7: store_cols<EPI_BIAS_W7>(e, acc, cs, c0, T); break;
            default: store_cols<EPI_NONE>(e, acc, cs, c0, T); break;
        }
    }
}

// 32 weights of one ggml block (or 32 f16 values) -> 4 x 16 bytes of fp16, element order 0..31
template <int TYPE> struct BlockRegs { uint32_t w[TYPE == DT_F16 ? 16 : (TYPE == DT_Q8_0 ? 9 : 6)]; };

// block `blk` of a raw row in shared memory (32-bit shared address) -> registers
template <int TYPE> __device__ __forceinline__ void read_block(uint32_t row, int blk, BlockRegs<TYPE> & r) {
    if constexpr (TYPE == DT_F16) {
#pragma unroll
        for (int i = 0; i < 4; i++) { const uint4 v = lds128(row + blk * 64 + i * 16); r.w[4 * i] = v.x; r.w[4 * i + 1] = v.y; r.w[4 * i + 2] = v.z; r.w[4 * i + 3] = v.w; }
    } else if constexpr (TYPE == DT_Q5_1) {          // 24-byte blocks on an 8-byte grid
#pragma unroll
        for (int i = 0; i < 3; i++) { const uint2 v = lds64(row + blk * 24 + i * 8); r.w[2 * i] = v.x; r.w[2 * i + 1] = v.y; }
    } else if constexpr (TYPE == DT_Q4_1) {          // 20-byte blocks on a 4-byte grid
#pragma unroll
        for (int i = 0; i < 5; i++) r.w[i] = lds32(row + blk * 20 + i * 4);
    } else {
        // 18 / 22 / 34-byte blocks start on a 2-byte grid: read whole words from the aligned-down address and realign
        constexpr int BB = QTraits<TYPE>::BLOCK_BYTES;
        constexpr int NW = BB / 4 + 1;
        const uint32_t start = (uint32_t) blk * BB;
        const uint32_t base = row + (start & ~3u);
        uint32_t t[NW];
#pragma unroll
        for (int i = 0; i < NW; i++) t[i] = lds32(base + i * 4);
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
    int raw_stages;                  // 2 or 3
    int raw_stage_bytes;             // ring slot size: the largest 128-row chunk of the batch's formats, 128-byte multiple
    int b_stages;                    // A/B ring depth, 2 .. MAX_STAGES
    const __half * act16[GEMV_MAX_PROBLEMS];   // canonical-layout fp16 activations per problem (convert_f16_kernel)
    const float * colscale[GEMV_MAX_PROBLEMS]; // [npad] power-of-two factor per token that the epilogue multiplies back in
    GemvProblem p[GEMV_MAX_PROBLEMS];          // first_cta / n_cta: CTA range of the problem = splits x tpad
    // Work decomposition. A problem's CTAs are [split][tile padded to a multiple of the cluster size]; the CTAs of one cluster are
    // cs neighbouring tiles of one split (tiles >= `tiles` are stand-ins that only take part in the B multicast). With splits > 1
    // every CTA contracts `steps_per_split` K-steps and leaves its 128 x npad partial tile in `partial`; the CTA that arrives last
    // at the tile's counter adds the partials up in split order (a fixed order: deterministic) and runs the epilogue.
    int cs;                                    // cluster size: 1, 2 or 4
    int tiles[GEMV_MAX_PROBLEMS], tpad[GEMV_MAX_PROBLEMS], splits[GEMV_MAX_PROBLEMS], steps_per_split[GEMV_MAX_PROBLEMS];
    int slot0[GEMV_MAX_PROBLEMS];              // first partial slot / counter of the problem (slot = slot0 + tile * splits + split)
    float * partial;                           // [slots][128][npad] fp32
    int * counters;                            // one per (problem, tile) with splits > 1, at index slot0 + tile * splits; zero between launches
    TraceRec * trace;
};

// What one CTA does: rows [tile * 128, +128) of problem `pi` against all tokens over K-steps [ks0, ks1).
struct TcWork { int pi, tile, split, ks0, ks1; bool standin; };

template <int TYPE>
__device__ void tc_tile(TcShared & sh, uint8_t * smem, const TcBatch & batch, const TcWork w) {
    using RT = RawTraits<TYPE>;
    const GemvProblem & P = sh.P;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NPAD = batch.npad, NG = NPAD / 8;
    const int row0 = w.tile * TILE_M;
    const uint32_t b_bytes = (uint32_t) NPAD * KSTEP * 2;
    constexpr uint32_t raw_bytes = (uint32_t) TILE_M * RT::ROW_STRIDE;
    uint8_t * const b_base = smem;
    const uint32_t tmem_a0 = sh.tmem_base + (uint32_t) NPAD;      // + stage * 32 columns
    const int nb = batch.b_stages;
    uint8_t * const raw_base = b_base + (size_t) nb * b_bytes;
    const uint32_t raw_slot = (uint32_t) batch.raw_stage_bytes;
    const int nsteps_total = P.K / KSTEP;
    const int nraw = batch.raw_stages;
    const int CS = batch.cs;
    const uint32_t rank = CS > 1 ? cluster_ctarank() : 0u;
    const uint16_t all_mask = (uint16_t) ((1u << CS) - 1u);
    const __half * act16 = batch.act16[w.pi];
    const int nsplit = batch.splits[w.pi];

    if (warp == WARP_RAW) {
        if (lane == 0 && !w.standin) {
            const int nchunks_total = (nsteps_total + RT::CHUNK_STEPS - 1) / RT::CHUNK_STEPS;
            const int c0 = w.ks0 / RT::CHUNK_STEPS, c1 = (w.ks1 + RT::CHUNK_STEPS - 1) / RT::CHUNK_STEPS;     // split boundaries are chunk-aligned
            const uint8_t * wt = reinterpret_cast<const uint8_t *>(P.Wt);
            int rs = 0; uint32_t ph = 0;
            for (int c = c0; c < c1; c++) {
                mbar_wait(&sh.raw_empty[rs], ph ^ 1);
                mbar_expect_tx(&sh.raw_full[rs], raw_bytes);
                bulk_load(raw_base + (size_t) rs * raw_slot, wt + ((size_t) w.tile * nchunks_total + c) * raw_bytes, raw_bytes, &sh.raw_full[rs]);
                if (++rs == nraw) { rs = 0; ph ^= 1; }
            }
        }
    } else if (warp == WARP_B) {
        if (lane == 0) {
            pdl_prologue();     // the activations come from the previous kernels
            const uint32_t slice = b_bytes / (uint32_t) CS;
            int s = 0; uint32_t ph = 0;
            for (int ks = w.ks0; ks < w.ks1; ks++) {
                mbar_wait(&sh.ab_empty[s], ph ^ 1);        // every CTA of the cluster has consumed this stage
                mbar_expect_tx(&sh.b_full[s], b_bytes);    // my barrier sees the whole stage: CS slices, one from each member
                const uint8_t * src = reinterpret_cast<const uint8_t *>(act16 + (size_t) ks * KSTEP * NPAD);
                if (CS == 1) bulk_load(b_base + (size_t) s * b_bytes, src, b_bytes, &sh.b_full[s]);
                else bulk_load_multicast(b_base + (size_t) s * b_bytes + (size_t) rank * slice, src + (size_t) rank * slice, slice, &sh.b_full[s], all_mask);
                if (++s == nb) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == WARP_MMA) {
        if (lane == 0 && w.standin) {
            // a stand-in has no rows: it releases every stage as soon as the stage's data has landed in its shared memory
            int s = 0; uint32_t ph = 0;
            for (int ks = w.ks0; ks < w.ks1; ks++) {
                mbar_wait(&sh.b_full[s], ph);
                for (int c = 0; c < CS; c++) mbar_arrive_remote(&sh.ab_empty[s], (uint32_t) c);
                if (++s == nb) { s = 0; ph ^= 1; }
            }
        } else if (lane == 0) {
            const uint32_t idesc = make_idesc(TILE_M, NPAD);
            // everything the issue loop needs sits in registers: TMEM addresses, the constant descriptor half and the 14-bit
            // start-address field of each B stage; per K-step the thread does two waits, 4 MMAs and ONE commit
            const uint32_t tmem_d = sh.tmem_base;
            const uint64_t desc_fixed = make_desc(0, (uint32_t) NG * 128, 128);
            const uint32_t b_addr0 = smem_u32(b_base) >> 4, b_stage16 = b_bytes >> 4, b_k16 = (uint32_t) (2 * NG * 128) >> 4;
            int s = 0; uint32_t ph = 0;
            for (int ks = w.ks0; ks < w.ks1; ks++) {
                const uint32_t a_col = tmem_a0 + (uint32_t) (s * 32);
                const uint64_t d0 = desc_fixed | (uint64_t) ((b_addr0 + (uint32_t) s * b_stage16) & 0x3FFFu);
                const uint64_t d1 = desc_fixed | (uint64_t) ((b_addr0 + (uint32_t) s * b_stage16 + b_k16) & 0x3FFFu);
                const uint64_t d2 = desc_fixed | (uint64_t) ((b_addr0 + (uint32_t) s * b_stage16 + 2 * b_k16) & 0x3FFFu);
                const uint64_t d3 = desc_fixed | (uint64_t) ((b_addr0 + (uint32_t) s * b_stage16 + 3 * b_k16) & 0x3FFFu);
                mbar_wait(&sh.b_full[s], ph);
                mbar_wait(&sh.a_full[s], ph);
                tc_fence_after_sync();
                umma_f16_ts(tmem_d, a_col, d0, idesc, ks > w.ks0 ? 1u : 0u);
                umma_f16_ts(tmem_d, a_col + 8, d1, idesc, 1u);
                umma_f16_ts(tmem_d, a_col + 16, d2, idesc, 1u);
                umma_f16_ts(tmem_d, a_col + 24, d3, idesc, 1u);
                // the A stage (mine) and the B stage (everybody's copy is written by everybody) are free once these MMAs have read them
                if (CS == 1) umma_commit(&sh.ab_empty[s]); else umma_commit_multicast(&sh.ab_empty[s], all_mask);
                if (++s == nb) { s = 0; ph ^= 1; }
            }
            umma_commit(&sh.acc_done);                 // ... and the accumulator is final
        }
    } else if (!w.standin) {
        // transform: thread = row r of the tile, BOTH blocks of a K-step (two independent dequantisation chains in
        // flight, one wait / store / arrive per 64 k); warps 0-3 take the even K-steps, warps 4-7 the odd ones, so two
        // A stages are being filled at any time. Warp w owns TMEM lanes 32 * (w % 4) .. + 31 = its 32 rows.
        const int r = tid & (TILE_M - 1), grp = tid >> 7;
        const uint32_t raw_row0 = smem_u32(raw_base) + (uint32_t) r * RT::ROW_STRIDE;
        const uint32_t tmem_a_mine = tmem_a0 + ((uint32_t) ((warp & 3) * 32) << 16);
        // trace marks of CTA 0, thread 0 (cycles, stored as start + cycles): [0] waiting for raw chunks, [1] reading + dequantising
        // two blocks, [2] waiting for a free A stage, [3] tcgen05.st + fences + arrive
        const bool acct = batch.trace != nullptr && blockIdx.x == 0 && tid == 0;
        long long c0 = 0, c1 = 0, c2 = 0, c3 = 0, tq = acct ? clock64() : 0;
        auto tick = [&](long long & a) { if (acct) { const long long now = clock64(); a += now - tq; tq = now; } };
        int s = grp % nb, rs = 0, cur_chunk = -1; uint32_t ph = 0, rph = 0;
        if (grp >= nb) ph ^= 1;        // (nb >= 2 always)
        for (int ks = w.ks0 + grp; ks < w.ks1; ks += 2) {
            const int c = ks / RT::CHUNK_STEPS, sc = ks % RT::CHUNK_STEPS;
            if (c != cur_chunk) {
                if (cur_chunk >= 0 && ++rs == nraw) { rs = 0; rph ^= 1; }
                cur_chunk = c;
                mbar_wait(&sh.raw_full[rs], rph);
            }
            tick(c0);
            BlockRegs<TYPE> regs0, regs1;
            read_block<TYPE>(raw_row0 + (uint32_t) rs * raw_slot, sc * 2, regs0);
            read_block<TYPE>(raw_row0 + (uint32_t) rs * raw_slot, sc * 2 + 1, regs1);
            // last step of this warp inside the chunk: the raw rows are in registers, hand the slot back
            const bool last_in_chunk = sc + 2 >= RT::CHUNK_STEPS || ks + 2 >= w.ks1;
            uint4 h0[4], h1[4];
            block_to_half<TYPE>(regs0, h0);
            block_to_half<TYPE>(regs1, h1);
            if (last_in_chunk) { __syncwarp(); if (lane == 0) mbar_arrive(&sh.raw_empty[rs]); }
            tick(c1);
            mbar_wait(&sh.ab_empty[s], ph ^ 1);
            tick(c2);
            tc_fence_after_sync();
            {
                uint32_t w32[32];
#pragma unroll
                for (int cc = 0; cc < 4; cc++) {
                    w32[4 * cc] = h0[cc].x; w32[4 * cc + 1] = h0[cc].y; w32[4 * cc + 2] = h0[cc].z; w32[4 * cc + 3] = h0[cc].w;
                    w32[16 + 4 * cc] = h1[cc].x; w32[16 + 4 * cc + 1] = h1[cc].y; w32[16 + 4 * cc + 2] = h1[cc].z; w32[16 + 4 * cc + 3] = h1[cc].w;
                }
                tmem_st32(tmem_a_mine + (uint32_t) (s * 32), w32);
            }
            tc_fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.a_full[s]);
            tick(c3);
            s += 2;
            if (s >= nb) { s -= nb; ph ^= 1; }
        }
        if (acct) {
            TraceRec * t = batch.trace;
            t->mark[0] = t->start + (unsigned long long) c0; t->mark[1] = t->start + (unsigned long long) c1;
            t->mark[2] = t->start + (unsigned long long) c2; t->mark[3] = t->start + (unsigned long long) c3;
        }
        {
            pdl_prologue();     // residual / gate inputs (and, with K-splits, the partial buffer) belong to the previous kernels until here
            if (tid < NPAD) sh.colscale[tid] = batch.colscale[w.pi][tid];
            asm volatile("bar.sync 2, 256;" ::: "memory");      // the 8 transform / epilogue warps
            mbar_wait(&sh.acc_done, 0);
            tc_fence_after_sync();
            if (nsplit == 1) {
                tc_epilogue_rows<0>(P, sh.colscale, sh.tmem_base, row0, NPAD, batch.T, nullptr, 1, 0);
            } else {
                const size_t slot_floats = (size_t) TILE_M * (size_t) ((NPAD + 31) & ~31);
                const int slot_first = batch.slot0[w.pi] + w.tile * nsplit;
                float * part0 = batch.partial + (size_t) slot_first * slot_floats;
                tc_epilogue_rows<1>(P, sh.colscale, sh.tmem_base, row0, NPAD, batch.T, part0 + (size_t) w.split * slot_floats, nsplit, slot_floats);
                __threadfence();                                    // my partial is visible device-wide before I take a ticket
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (tid == 0) sh.split_rank = atomicAdd(batch.counters + slot_first, 1);
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (sh.split_rank == nsplit - 1) {                  // every other split's partial has been published before its ticket
                    __threadfence();
                    tc_epilogue_rows<2>(P, sh.colscale, sh.tmem_base, row0, NPAD, batch.T, part0, nsplit, slot_floats);
                    if (tid == 0) batch.counters[slot_first] = 0;   // ready for the next launch (ordered by kernel completion)
                }
            }
        }
    } else {
        pdl_prologue();
    }
    tc_fence_before_sync();
    __syncthreads();
}

__global__ void __launch_bounds__(THREADS, 1) gemm_tc_kernel(const TcBatch batch) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ TcShared sh;
    trace_begin(batch.trace);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    int pi = 0;
    for (int i = 1; i < batch.n; i++) if ((int) blockIdx.x >= batch.p[i].first_cta) pi = i;
    if (threadIdx.x == 0) {
        sh.t0 = clock64();
        sh.P = batch.p[pi];
        for (int s = 0; s < MAX_RAW_STAGES; s++) { mbar_init(&sh.raw_full[s], 1); mbar_init(&sh.raw_empty[s], XFORM_WARPS); }
        for (int s = 0; s < MAX_STAGES; s++) { mbar_init(&sh.a_full[s], XFORM_WARPS / 2); mbar_init(&sh.b_full[s], 1); mbar_init(&sh.ab_empty[s], (uint32_t) batch.cs); }
        mbar_init(&sh.acc_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(&sh.tmem_base, (uint32_t) batch.tmem_cols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    if (batch.cs > 1) cluster_sync_all();          // nobody multicasts into, or arrives on, a barrier that is not initialised yet
    TcWork w;
    w.pi = pi;
    {
        const int local = (int) blockIdx.x - sh.P.first_cta, tpad = batch.tpad[pi];
        w.split = local / tpad;
        w.tile = local % tpad;
        w.standin = w.tile >= batch.tiles[pi];
        const int nsteps = sh.P.K / KSTEP, per = batch.steps_per_split[pi];
        w.ks0 = w.split * per;
        w.ks1 = min(nsteps, w.ks0 + per);
    }
    switch (sh.P.type) {
        case DT_Q4_0: tc_tile<DT_Q4_0>(sh, smem, batch, w); break;
        case DT_Q4_1: tc_tile<DT_Q4_1>(sh, smem, batch, w); break;
        case DT_Q5_0: tc_tile<DT_Q5_0>(sh, smem, batch, w); break;
        case DT_Q5_1: tc_tile<DT_Q5_1>(sh, smem, batch, w); break;
        case DT_Q8_0: tc_tile<DT_Q8_0>(sh, smem, batch, w); break;
        default: tc_tile<DT_F16>(sh, smem, batch, w); break;
    }
    if (batch.cs > 1) cluster_sync_all();          // my shared memory and barriers stay valid until every member is done with them
    if (threadIdx.x < 32) tmem_dealloc(sh.tmem_base, (uint32_t) batch.tmem_cols);
    trace_end(batch.trace);
}

// x fp32 [K, T] column-major (column t contiguous) -> the fp16 B operand in the UMMA canonical layout, one contiguous B stage per
// K-step:  [K / 64][kc = 8][g = npad / 8][8 tokens][8 k]  (tokens >= T zero-filled). One CTA = one token column of one distinct
// input of the GEMM batch (blockIdx.y = input; up to 8: x_r, x_k, x_v, x_g, x_w of the time mix).
//
// What the halves hold. For QUANTISED weights the reference multiplies Q8_0 / Q8_1 activation blocks (ggml-cpu.c:7439-7458,
// quantize_row_q8_0 / q8_1 ggml-cpu-quants.c:781-846, 1085-1160): per 32 elements d = fp16(amax / 127), q = rint(x * 127 / amax),
// and the dot product sees d * q. So does this path: the half written is d * q (exact in fp32: 11 x 7 bits) rounded once to fp16,
// i.e. the tensor cores contract the reference's own operand values instead of a "more accurate" fp16 rounding of x, and the
// >= 32-token path tracks the dp4a path to fp16 rounding of the products (2^-12 relative per element) instead of to the Q8
// quantisation noise (~4e-3 of the block maximum). For F16 weights the reference rounds activations to fp16
// (ggml-cpu.c:259-264): the half is fp16(x).
// Range. fp16 ends at 65504 where the reference's operands do not (d is fp16 but q * d goes up to amax): every token column
// carries a power-of-two scale 2^-e, e = max(0, exponent(column maximum) - 15), applied before the rounding (exact) and taken
// out by the GEMM epilogue through colscale[t] = 2^e, so activations up to 3e38 / 127 stay finite and equally precise.
struct ConvertBatch {
    int n, T, npad;
    const float * x[GEMV_MAX_PROBLEMS];
    long long ldx[GEMV_MAX_PROBLEMS];
    int K[GEMV_MAX_PROBLEMS];
    int quant[GEMV_MAX_PROBLEMS];      // 1: Q8 block values (quantised weights), 0: plain fp16 rounding (F16 weights)
    __half * out[GEMV_MAX_PROBLEMS];
    float * colscale[GEMV_MAX_PROBLEMS];   // [npad] per input
    TraceRec * trace;
};
constexpr int CVT_THREADS = 256;
__global__ void __launch_bounds__(CVT_THREADS) convert_f16_kernel(const ConvertBatch cb) {
    __shared__ float red[CVT_THREADS / 32];
    trace_begin(cb.trace);
    pdl_prologue();
    const int q = blockIdx.y, t = blockIdx.x, T = cb.T, npad = cb.npad, K = cb.K[q];
    const int tid = threadIdx.x;
    uint4 * out = reinterpret_cast<uint4 *>(cb.out[q]);
    const int NG = npad / 8, g = t >> 3, t8 = t & 7;
    const int nblk = K / 32;
    // 16-byte chunk (ks, kc) of this token sits at ((ks * 8 + kc) * NG + g) * 8 + t8
    if (t >= T) {
        for (int c = tid; c < K / 8; c += CVT_THREADS) out[((size_t) c * NG + g) * 8 + t8] = make_uint4(0, 0, 0, 0);
        if (tid == 0) cb.colscale[q][t] = 1.0f;
        return;
    }
    const float4 * x4 = reinterpret_cast<const float4 *>(cb.x[q] + (long long) t * cb.ldx[q]);
    float cmax = 0.f;
    for (int i = tid; i < K / 4; i += CVT_THREADS) {
        const float4 v = x4[i];
        cmax = fmaxf(cmax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, o));
    if ((tid & 31) == 0) red[tid >> 5] = cmax;
    __syncthreads();
    cmax = red[0];
#pragma unroll
    for (int i = 1; i < CVT_THREADS / 32; i++) cmax = fmaxf(cmax, red[i]);
    int ex = 0;
    if (cmax > 32768.0f && cmax <= 3.0e38f) { (void) frexpf(cmax, &ex); ex -= 15; }      // cmax <= 2^ex_raw  ->  cmax * 2^-ex <= 2^15
    const float down = ex > 0 ? exp2f((float) -ex) : 1.0f;
    if (tid == 0) cb.colscale[q][t] = ex > 0 ? exp2f((float) ex) : 1.0f;
    const bool quant = cb.quant[q] != 0;
    for (int blk = tid; blk < nblk; blk += CVT_THREADS) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = x4[blk * 8 + j];      // second touch of the column: L1 / L2
        if (quant) {
            float amax = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
            const float d = __half2float(__float2half_rn(amax / 127.0f)) * down;
            const float id = (amax != 0.0f) ? 127.0f / amax : 0.0f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                v[j].x = (float) __float2int_rn(v[j].x * id) * d; v[j].y = (float) __float2int_rn(v[j].y * id) * d;
                v[j].z = (float) __float2int_rn(v[j].z * id) * d; v[j].w = (float) __float2int_rn(v[j].w * id) * d;
            }
        } else if (ex > 0) {
#pragma unroll
            for (int j = 0; j < 8; j++) { v[j].x *= down; v[j].y *= down; v[j].z *= down; v[j].w *= down; }
        }
        const int c0 = blk * 4;            // first of the block's four 8-element chunks
#pragma unroll
        for (int cc = 0; cc < 4; cc++) {
            __half2 a = __floats2half2_rn(v[2 * cc].x, v[2 * cc].y), b = __floats2half2_rn(v[2 * cc].z, v[2 * cc].w);
            __half2 c = __floats2half2_rn(v[2 * cc + 1].x, v[2 * cc + 1].y), d2 = __floats2half2_rn(v[2 * cc + 1].z, v[2 * cc + 1].w);
            out[((size_t) (c0 + cc) * NG + g) * 8 + t8] =
                make_uint4(*reinterpret_cast<uint32_t *>(&a), *reinterpret_cast<uint32_t *>(&b), *reinterpret_cast<uint32_t *>(&c), *reinterpret_cast<uint32_t *>(&d2));
        }
    }
    trace_end(cb.trace);
}

// Row-major matrix (rows of `pitch` bytes) -> tile-major prefill copy. One thread per 4-byte word of the destination.
__global__ void tc_repack_kernel(const uint8_t * src, long long pitch, int M, int nchunks, int chunk_bytes, int row_stride, uint8_t * dst) {
    const int words_per_row = row_stride / 4;
    const long long stage_words = (long long) TILE_M * words_per_row;
    const long long total = (long long) gridDim.y * nchunks * stage_words;      // gridDim.y = tiles
    for (long long w = (long long) blockIdx.x * blockDim.x + threadIdx.x + (long long) blockIdx.y * nchunks * stage_words;
         w < ((long long) blockIdx.y + 1) * nchunks * stage_words && w < total; w += (long long) gridDim.x * blockDim.x) {
        const long long in_tile = w - (long long) blockIdx.y * nchunks * stage_words;
        const int c = (int) (in_tile / stage_words);
        const int rem = (int) (in_tile % stage_words);
        const int r = rem / words_per_row, off = (rem % words_per_row) * 4;
        const int row = blockIdx.y * TILE_M + r;
        const long long col = (long long) c * chunk_bytes + off;
        uint32_t v = 0;
        if (row < M && off < chunk_bytes && col + 4 <= pitch) v = *reinterpret_cast<const uint32_t *>(src + (long long) row * pitch + col);
        reinterpret_cast<uint32_t *>(dst)[w] = v;
    }
}

}  // namespace tc

bool gemm_tc_eligible(int type, int K) { return type != DT_F32 && K % tc::KSTEP == 0 && K >= tc::KSTEP; }

size_t gemm_tc_tiled_bytes(int type, int M, int K) {
    if (!gemm_tc_eligible(type, K)) return 0;
    const tc::RawGeom g = tc::raw_geom(type);
    const size_t nsteps = (size_t) K / tc::KSTEP, nchunks = (nsteps + g.chunk_steps - 1) / g.chunk_steps, ntiles = ((size_t) M + tc::TILE_M - 1) / tc::TILE_M;
    return ntiles * nchunks * tc::TILE_M * (size_t) g.row_stride;
}

cudaError_t gemm_tc_repack(const void * W, long long pitch, int type, int M, int K, void * dst, cudaStream_t stream) {
    if (!gemm_tc_eligible(type, K)) return cudaErrorInvalidValue;
    const tc::RawGeom g = tc::raw_geom(type);
    const int nsteps = K / tc::KSTEP, nchunks = (nsteps + g.chunk_steps - 1) / g.chunk_steps, ntiles = (M + tc::TILE_M - 1) / tc::TILE_M;
    const long long words_per_tile = (long long) nchunks * tc::TILE_M * (g.row_stride / 4);
    int bx = (int) ((words_per_tile + 255) / 256);
    if (bx > 64) bx = 64;
    tc::tc_repack_kernel<<<dim3(bx, ntiles), 256, 0, stream>>>(reinterpret_cast<const uint8_t *>(W), pitch, M, nchunks, g.chunk_bytes, g.row_stride,
                                                               reinterpret_cast<uint8_t *>(dst));
    return cudaGetLastError();
}

bool gemm_tc_supported(const GemvProblem & p, int T) {
    return T >= 16 && T <= tc::MAX_N && p.Wt != nullptr && gemm_tc_eligible(p.type, p.K) && (p.ldx % 4) == 0 &&
           (reinterpret_cast<uintptr_t>(p.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.Wt) & 15) == 0;
}

// Workspace layout (gemm_tc_workspace_bytes): [split-K tile counters, zero between launches][split-K partial tiles][fp16 operands +
// per-token scales of the batch's distinct inputs]. The caller zeroes the counter block once after allocating.
size_t gemm_tc_workspace_bytes(int T, size_t operand_halves) {
    const size_t npad = (size_t) (T + 15) / 16 * 16;
    return GEMM_TC_COUNTER_BYTES + GEMM_TC_PARTIAL_BYTES + operand_halves * 2 + (size_t) GEMV_MAX_PROBLEMS * (npad * 4 + 512);
}

cudaError_t gemm_tc_launch(GemvBatch & batch, const DeviceInfo & dev, cudaStream_t stream, void * workspace, size_t workspace_bytes) {
    if (workspace_bytes < GEMM_TC_COUNTER_BYTES + GEMM_TC_PARTIAL_BYTES) return cudaErrorMemoryAllocation;
    tc::TcBatch tb;
    memset(&tb, 0, sizeof(tb));
    tb.n = batch.n; tb.T = batch.T;
    tb.npad = (batch.T + 15) / 16 * 16;
    tb.tmem_cols = 32;
    while (tb.tmem_cols < tb.npad + tc::MAX_STAGES * 32) tb.tmem_cols *= 2;      // accumulator + A stages
    tb.counters = reinterpret_cast<int *>(workspace);
    tb.partial = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(workspace) + GEMM_TC_COUNTER_BYTES);

    uint8_t * scratch = reinterpret_cast<uint8_t *>(workspace) + GEMM_TC_COUNTER_BYTES + GEMM_TC_PARTIAL_BYTES;
    const size_t scratch_bytes = workspace_bytes - GEMM_TC_COUNTER_BYTES - GEMM_TC_PARTIAL_BYTES;
    size_t used = 0;
    tc::ConvertBatch cvt;
    memset(&cvt, 0, sizeof(cvt));
    cvt.T = batch.T; cvt.npad = tb.npad;
    int share_of[GEMV_MAX_PROBLEMS];
    for (int i = 0; i < batch.n; i++) {
        GemvProblem & p = batch.p[i];
        const int quant = p.type != DT_F16;
        const size_t need = (((size_t) tb.npad * p.K * sizeof(__half) + 255) & ~(size_t) 255) + (((size_t) tb.npad * sizeof(float) + 255) & ~(size_t) 255);
        // problems that share an input (and its operand kind) share the converted copy
        int shared = -1;
        for (int j = 0; j < i; j++)
            if (batch.p[j].x == p.x && batch.p[j].K == p.K && batch.p[j].ldx == p.ldx && (batch.p[j].type != DT_F16) == (quant != 0)) shared = share_of[j];
        if (shared < 0) {
            if (used + need > scratch_bytes) return cudaErrorMemoryAllocation;
            shared = cvt.n;
            cvt.x[cvt.n] = p.x; cvt.ldx[cvt.n] = p.ldx; cvt.K[cvt.n] = p.K; cvt.quant[cvt.n] = quant;
            cvt.out[cvt.n] = reinterpret_cast<__half *>(scratch + used);
            cvt.colscale[cvt.n] = reinterpret_cast<float *>(scratch + used + (((size_t) tb.npad * p.K * sizeof(__half) + 255) & ~(size_t) 255));
            cvt.n++;
            used += need;
        }
        share_of[i] = shared;
        tb.act16[i] = cvt.out[shared];
        tb.colscale[i] = cvt.colscale[shared];
    }
    // ---- work decomposition: cluster size, K-splits ----
    // Clusters of 4 neighbouring row tiles share every B stage through multicast: a CTA pulls 1/4 of the activations it multiplies
    // out of L2 (the B operand, re-read by every row tile, is 3x the weight bytes of a 128-row tile per K-step). Launches that
    // would leave most of the 148 SMs idle (matrices with 4096 rows: 32 tiles; the LoRA matrices: 1-2 tiles) are cut along K.
    static const int force_cs = [] { const char * e = getenv("RWKV_B200_TC_CLUSTER"); return e ? atoi(e) : 0; }();
    static const int force_split = [] { const char * e = getenv("RWKV_B200_TC_SPLITK"); return e ? atoi(e) : -1; }();
    int base = 0;
    for (int i = 0; i < batch.n; i++) { tb.tiles[i] = (batch.p[i].M + tc::TILE_M - 1) / tc::TILE_M; base += tb.tiles[i]; }
    tb.cs = base >= 16 ? 4 : 1;
    if (force_cs == 1 || force_cs == 2 || force_cs == 4) tb.cs = force_cs;
    int total = 0;
    for (int i = 0; i < batch.n; i++) { tb.tpad[i] = (tb.tiles[i] + tb.cs - 1) / tb.cs * tb.cs; total += tb.tpad[i]; }
    const int sms = dev.num_sms > 0 ? dev.num_sms : 148;
    int want = total * 5 < sms * 3 ? sms / total : 1;      // below 60 % of the SMs: split
    if (force_split >= 0) want = force_split < 1 ? 1 : force_split;
    const size_t slot_bytes = (size_t) tc::TILE_M * (size_t) ((tb.npad + 31) & ~31) * sizeof(float);
    const int slot_cap = (int) (GEMM_TC_PARTIAL_BYTES / slot_bytes), ctr_cap = (int) (GEMM_TC_COUNTER_BYTES / sizeof(int));
    int next = 0, slots = 0;
    for (int i = 0; i < batch.n; i++) {
        GemvProblem & p = batch.p[i];
        const int nsteps = p.K / tc::KSTEP, chunk = tc::raw_geom(p.type).chunk_steps;
        int splits = want;
        if (splits > 32) splits = 32;
        const int max_splits = (nsteps + chunk - 1) / chunk;          // at least one raw chunk per split
        if (splits > max_splits) splits = max_splits;
        if (splits < 1) splits = 1;
        int per = (nsteps + splits - 1) / splits;
        per = (per + chunk - 1) / chunk * chunk;                       // split boundaries on raw-chunk boundaries
        splits = (nsteps + per - 1) / per;
        if (splits > 1 && (slots + tb.tiles[i] * splits > slot_cap || slots + tb.tiles[i] * splits > ctr_cap)) { splits = 1; per = nsteps; }
        tb.splits[i] = splits;
        tb.steps_per_split[i] = splits == 1 ? nsteps : per;
        tb.slot0[i] = slots;
        if (splits > 1) slots += tb.tiles[i] * splits;
        p.first_cta = next;
        p.n_cta = tb.tpad[i] * splits;
        next += p.n_cta;
        tb.p[i] = p;
    }
    {   // every distinct input of the batch -> fp16 canonical layout, one launch: one CTA per (token column, input)
        cvt.trace = trace_slot("convert_f16");
        g_kernel_launches++;
        cudaError_t e = launch_pdl(tc::convert_f16_kernel, dim3(tb.npad, cvt.n), dim3(tc::CVT_THREADS), 0, stream, cvt);
        if (e != cudaSuccess) return e;
    }
    // shared memory: 3 raw chunks (sized for the widest format of this batch) and as many B stages as fit, at most 8. The ring has
    // to be deep: a stage comes back only after store -> arrive -> MMA issue -> MMA -> commit -> wake-up, and each of the two
    // transform groups owns every other stage
    constexpr size_t smem_budget = 227 * 1024 - 2048;
    size_t raw_slot = 0;
    for (int i = 0; i < batch.n; i++) {
        const size_t b = (size_t) tc::TILE_M * tc::raw_geom(batch.p[i].type).row_stride;
        if (b > raw_slot) raw_slot = b;
    }
    raw_slot = (raw_slot + 127) & ~(size_t) 127;
    tb.raw_stages = tc::MAX_RAW_STAGES;
    tb.raw_stage_bytes = (int) raw_slot;
    const size_t fixed = (size_t) tb.raw_stages * raw_slot;
    const size_t b_bytes = (size_t) tb.npad * tc::KSTEP * 2;
    int nb = (int) ((smem_budget - fixed) / b_bytes);
    if (nb > tc::MAX_STAGES) nb = tc::MAX_STAGES;
    if (nb < 2) return cudaErrorInvalidValue;
    tb.b_stages = nb;
    const size_t smem = fixed + (size_t) nb * b_bytes;
    static PerDeviceOnce once;               // the opt-in is per device
    const cudaError_t ae = once.run([&] { return cudaFuncSetAttribute(tc::gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem_budget); });
    if (ae != cudaSuccess) return ae;
    tb.trace = trace_slot("gemm_tc");
    g_kernel_launches++;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned) next);
    cfg.blockDim = dim3(tc::THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (g_use_pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; na++; }
    if (tb.cs > 1) { attr[na].id = cudaLaunchAttributeClusterDimension; attr[na].val.clusterDim.x = (unsigned) tb.cs; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; na++; }
    cfg.attrs = attr;
    cfg.numAttrs = (unsigned) na;
    return cudaLaunchKernelEx(&cfg, tc::gemm_tc_kernel, tb);
}

}  // namespace rwkv
