// The staged activation column: the operand a single-token GEMV multiplies, in the exact byte layout the streaming kernel
// (gemv_tma_device.cuh) keeps in shared memory. Producer kernels that own whole 32-element blocks of a vector (LayerNorm + mix,
// the v6 lerp, the WKV kernels) emit it next to their fp32 output, so the consuming GEMV's staging step shrinks from "load 16 KB of
// fp32, reduce, divide, round, pack" (4-7 us on the critical path of EVERY launch, profiles/r2_trace_decode_c3.csv) to one 5 KB copy.
//
// What the column holds is what the reference's CPU path multiplies with (ggml-cpu.c:253-311 vec_dot_type): Q8_0 blocks for
// Q4_0 / Q5_0 / Q8_0 weights, Q8_1 blocks for Q4_1 / Q5_1 (x86 flavour of quantize_row_q8_0 / q8_1, ggml-cpu-quants.c:781-846,
// 1085-1160: d = fp16(amax / 127), q = rint(x * 127 / amax), s = fp16(d32 * sum q)), fp16 for F16 weights, fp32 for F32 weights.
// The block maximum and the integer block sum do not depend on the order they are taken in, so whoever quantises a block produces
// the same bytes: the hand-off changes no result bit.
//
// Layout for block b = u * UB + bi (UB = 1 for the *_1 formats, 2 for the 2-byte-aligned *_0 formats), half h (elements 16h..16h+15):
//     q     at  ((bi * 2 + h) * nunits + u) * 16
//     scale at  nunits * UB * 32 + (bi * nunits + u) * 8          (ActScale {d, s})
#pragma once
#include "quant_decode.cuh"

#include <cuda_fp16.h>

namespace rwkv {
namespace act {

// Which staged format a weight type multiplies: weights of the same class share one staged column.
enum StageClass : int { SC_NONE = -1, SC_F32 = 0, SC_F16 = 1, SC_Q8_0 = 2, SC_Q8_1 = 3 };
__host__ __device__ inline int stage_class(int type) {
    switch (type) {
        case DT_F32: return SC_F32;
        case DT_F16: return SC_F16;
        case DT_Q4_0: case DT_Q5_0: case DT_Q8_0: return SC_Q8_0;
        case DT_Q4_1: case DT_Q5_1: return SC_Q8_1;
        default: return SC_NONE;
    }
}
// Blocks of a quantised activation column rounded up to whole units (pairs for the 2-byte-aligned formats).
__host__ __device__ inline int padded_blocks(int type, int K) {
    const int nblk = K / 32;
    return (type == DT_Q4_1 || type == DT_Q5_1) ? nblk : (nblk + 1) / 2 * 2;
}
__host__ __device__ inline size_t bytes_per_column(int type, int K) {
    size_t b;
    if (type == DT_F32) b = (size_t) K * 4;
    else if (type == DT_F16) b = (size_t) K * 2;
    else b = (size_t) padded_blocks(type, K) * (32 + sizeof(ActScale));
    return (b + 15) & ~(size_t) 15;
}

#if defined(__CUDACC__)
// A producer's view of one staged column it has to fill (dst == nullptr: nothing to emit).
struct StagedOut {
    uint8_t * dst;
    int type;      // weight type of the consumer (only its stage class matters)
    int K;         // length of the vector, a multiple of 32
};

// One full warp = one 32-element block: lane l holds element 32 * blk + l. All 32 lanes must call.
__device__ __forceinline__ void warp_emit_block(const StagedOut & o, int blk, float v) {
    const int lane = threadIdx.x & 31;
    const int cls = stage_class(o.type);
    if (cls == SC_F32) { reinterpret_cast<float *>(o.dst)[blk * 32 + lane] = v; return; }
    if (cls == SC_F16) { reinterpret_cast<__half *>(o.dst)[blk * 32 + lane] = __float2half_rn(v); return; }
    const bool has_min = cls == SC_Q8_1;
    const int UB = has_min ? 1 : 2;
    const int nblk = o.K / 32, nunits = (nblk + UB - 1) / UB;
    const float amax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fabsf(v))));   // non-negative floats order as integers
    const float d32 = amax / 127.0f;
    const float id = (amax != 0.0f) ? 127.0f / amax : 0.0f;
    const int q = __float2int_rn(v * id);
    const int isum = __reduce_add_sync(0xffffffffu, q);
    int word = q & 0xFF;
    word |= (__shfl_down_sync(0xffffffffu, q, 1) & 0xFF) << 8;
    word |= (__shfl_down_sync(0xffffffffu, q, 2) & 0xFF) << 16;
    word |= (__shfl_down_sync(0xffffffffu, q, 3) & 0xFF) << 24;
    const int u = blk / UB, bi = blk % UB;
    if ((lane & 3) == 0) {
        const int i8 = lane >> 2;                               // word i8 = elements 4 i8 .. 4 i8 + 3 of the block
        *reinterpret_cast<int *>(o.dst + ((size_t) (bi * 2 + (i8 >> 2)) * nunits + u) * 16 + (i8 & 3) * 4) = word;
    }
    if (lane == 0) {
        ActScale a;
        a.d = __half2float(__float2half_rn(d32));
        a.s = has_min ? __half2float(__float2half_rn(d32 * (float) isum)) : (float) isum;
        *reinterpret_cast<ActScale *>(o.dst + (size_t) nunits * UB * 32 + ((size_t) bi * nunits + u) * 8) = a;
    }
}
#endif

}  // namespace act
}  // namespace rwkv
