// Decoding of ggml quant blocks (ggml-common.h:161-221; element order ggml-quants.c:255-347) into
// int8x4 words ready for dp4a, shared by the decode-GEMV and the prefill GEMM.
//
// HBM layout = file layout: rows of 18/20/22/24/34-byte blocks, row pitch padded to 16 B -- with ONE difference: the 32-bit word
// of fifth bits (qh) of every Q5_0 / Q5_1 block is stored bit-transposed (qh5_to_device, applied once at load time; same bytes,
// same information). In the file, bit 4w + k of qh belongs to byte k of int8x4 word w; on the device it sits at bit 8k + w, so
// one shift + one AND drops the four fifth bits of a word onto bit 4 of its four bytes, where the file layout needs a shift, an
// AND, an integer multiply and another AND per word (spread_bit5). The Q5_1 decode shrinks from ~52 to ~27 instructions per block
// (the dp4a GEMV is instruction-bound on it: profiles/r2_c4 trace, 0.57 issue slots busy per cycle at 3.5 TB/s).
// A "unit" is the smallest group of whole blocks that is 4-byte aligned: 2 blocks for the
// 2-byte-aligned formats (Q4_0 36 B, Q5_0 44 B, Q8_0 68 B), 1 block for Q4_1 (20 B) / Q5_1 (24 B).
// A lane pulls one unit as 32-bit words and realigns with funnel shifts.
//
// Every function is __host__ __device__ so tests/host_kernels_check.cu can verify the bit twiddling
// on the CPU against a scalar restatement before any GPU time is spent.
#pragma once
#include <cmath>
#include <cstdint>
#include "../formats.h"

#if defined(__CUDACC__)
#  include <cuda_fp16.h>
#  define RWKV_HD __host__ __device__ __forceinline__
#else
#  define RWKV_HD inline
#endif

namespace rwkv {

RWKV_HD uint32_t funnel16(uint32_t lo, uint32_t hi) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, 16);
#else
    return (lo >> 16) | (hi << 16);
#endif
}

RWKV_HD int dot4_i8(int a, int b, int c) {   // c + sum of 4 signed-byte products
#if defined(__CUDA_ARCH__)
    return __dp4a(a, b, c);
#else
    for (int i = 0; i < 4; i++) c += (int) (int8_t) (a >> (8 * i)) * (int) (int8_t) (b >> (8 * i));
    return c;
#endif
}

RWKV_HD float half_bits_to_float(uint32_t h16) {
#if defined(__CUDA_ARCH__)
    return __half2float(__ushort_as_half((unsigned short) h16));
#else
    return fp16_to_fp32((uint16_t) h16);
#endif
}

// 4 bits b0..b3 of n -> bit 4 of bytes 0..3. The four shifted copies n<<4, n<<11, n<<18, n<<25 occupy
// disjoint bit ranges, so the multiply has no carries.
RWKV_HD uint32_t spread_bit5(uint32_t n4) { return (n4 * 0x02040810u) & 0x10101010u; }

// File order -> device order of a Q5 block's fifth-bit word (an 8 x 4 bit transpose) and back.
RWKV_HD uint32_t qh5_to_device(uint32_t qh) {
    uint32_t o = 0;
    for (int w = 0; w < 8; w++)
        for (int k = 0; k < 4; k++) o |= ((qh >> (4 * w + k)) & 1u) << (8 * k + w);
    return o;
}
RWKV_HD uint32_t qh5_to_file(uint32_t qd) {
    uint32_t o = 0;
    for (int w = 0; w < 8; w++)
        for (int k = 0; k < 4; k++) o |= ((qd >> (8 * k + w)) & 1u) << (4 * w + k);
    return o;
}

template <int TYPE> struct QTraits;
template <> struct QTraits<DT_Q4_0> { static constexpr int BLOCK_BYTES = 18, UNIT_BLOCKS = 2, UNIT_WORDS = 9,  OFFSET = 8,  HAS_MIN = 0; };
template <> struct QTraits<DT_Q4_1> { static constexpr int BLOCK_BYTES = 20, UNIT_BLOCKS = 1, UNIT_WORDS = 5,  OFFSET = 0,  HAS_MIN = 1; };
template <> struct QTraits<DT_Q5_0> { static constexpr int BLOCK_BYTES = 22, UNIT_BLOCKS = 2, UNIT_WORDS = 11, OFFSET = 16, HAS_MIN = 0; };
template <> struct QTraits<DT_Q5_1> { static constexpr int BLOCK_BYTES = 24, UNIT_BLOCKS = 1, UNIT_WORDS = 6,  OFFSET = 0,  HAS_MIN = 1; };
template <> struct QTraits<DT_Q8_0> { static constexpr int BLOCK_BYTES = 34, UNIT_BLOCKS = 2, UNIT_WORDS = 17, OFFSET = 0,  HAS_MIN = 0; };

// One decoded block: q[0..3] hold elements 0..15 (4 per word, element 4i+k in byte k of q[i]),
// q[4..7] elements 16..31. Values are the stored integers (0..15 / 0..31 / signed for Q8_0);
// the weight is (q - OFFSET) * d + m.
struct BlockQ {
    int q[8];
    float d, m;
};

// qh: the fifth bits in DEVICE order (qh5_to_device): word w's four bits at positions 8k + w.
RWKV_HD void nibble_words(const uint32_t qs[4], uint32_t qh, bool five_bit, int q[8]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint32_t lo = qs[i] & 0x0F0F0F0Fu, hi = (qs[i] >> 4) & 0x0F0F0F0Fu;
        if (five_bit) {
            lo |= (qh << (4 - i)) & 0x10101010u;        // word i:     bits 8k + i     -> 8k + 4
            hi |= (qh >> i) & 0x10101010u;              // word 4 + i: bits 8k + 4 + i -> 8k + 4
        }
        q[i] = (int) lo;
        q[4 + i] = (int) hi;
    }
}

// w = the UNIT_WORDS 32-bit words of one unit; b = block index inside the unit.
template <int TYPE> RWKV_HD void decode_block(const uint32_t * w, int b, BlockQ & o);

template <> RWKV_HD void decode_block<DT_Q4_0>(const uint32_t * w, int b, BlockQ & o) {
    uint32_t qs[4];
    if (b == 0) {
        o.d = half_bits_to_float(w[0] & 0xFFFFu);
#pragma unroll
        for (int i = 0; i < 4; i++) qs[i] = funnel16(w[i], w[i + 1]);
    } else {
        o.d = half_bits_to_float(w[4] >> 16);
#pragma unroll
        for (int i = 0; i < 4; i++) qs[i] = w[5 + i];
    }
    o.m = 0.0f;
    nibble_words(qs, 0, false, o.q);
}
template <> RWKV_HD void decode_block<DT_Q4_1>(const uint32_t * w, int, BlockQ & o) {
    o.d = half_bits_to_float(w[0] & 0xFFFFu);
    o.m = half_bits_to_float(w[0] >> 16);
    nibble_words(w + 1, 0, false, o.q);
}
template <> RWKV_HD void decode_block<DT_Q5_0>(const uint32_t * w, int b, BlockQ & o) {
    uint32_t qs[4], qh;
    if (b == 0) {
        o.d = half_bits_to_float(w[0] & 0xFFFFu);
        qh = funnel16(w[0], w[1]);
#pragma unroll
        for (int i = 0; i < 4; i++) qs[i] = funnel16(w[1 + i], w[2 + i]);
    } else {
        o.d = half_bits_to_float(w[5] >> 16);
        qh = w[6];
#pragma unroll
        for (int i = 0; i < 4; i++) qs[i] = w[7 + i];
    }
    o.m = 0.0f;
    nibble_words(qs, qh, true, o.q);
}
template <> RWKV_HD void decode_block<DT_Q5_1>(const uint32_t * w, int, BlockQ & o) {
    o.d = half_bits_to_float(w[0] & 0xFFFFu);
    o.m = half_bits_to_float(w[0] >> 16);
    nibble_words(w + 2, w[1], true, o.q);
}
template <> RWKV_HD void decode_block<DT_Q8_0>(const uint32_t * w, int b, BlockQ & o) {
    if (b == 0) {
        o.d = half_bits_to_float(w[0] & 0xFFFFu);
#pragma unroll
        for (int i = 0; i < 8; i++) o.q[i] = (int) funnel16(w[i], w[i + 1]);
    } else {
        o.d = half_bits_to_float(w[8] >> 16);
#pragma unroll
        for (int i = 0; i < 8; i++) o.q[i] = (int) w[9 + i];
    }
    o.m = 0.0f;
}

// Activation block as the reference's x86 path quantises it (ggml-cpu-quants.c:781-846 for Q8_0,
// :1085-1160 for Q8_1): d = fp16(amax/127), q = rint(x * 127/amax), and per block either
//   isum = sum(q)            (exact int, used to fold the -8 / -16 offset of Q4_0 / Q5_0), or
//   s    = fp16(d32 * sum(q)) (the Q8_1 `s` field that multiplies the block minimum of Q4_1 / Q5_1).
struct ActScale {
    float d;   // fp16-rounded scale, as float
    float s;   // HAS_MIN ? fp16(d32*sum q) : (float) sum(q)   (ints up to 32*127 are exact in fp32)
};

// Contribution of one weight block against one activation block, exactly the reference formulas
// (ggml-cpu-quants.c:2302-2316, 2594-2609, 2944-2963, 3318-3338, 3690-3698) with fp32 accumulation:
//   _0 types: acc += (dW*dA) * sum((q-OFFSET)*a)        _1 types: acc += (dW*dA) * sum(q*a) + mW*sA
template <int TYPE> RWKV_HD float block_dot(const BlockQ & wq, const int * a8, ActScale as, float acc) {
    int isum = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) isum = dot4_i8(wq.q[i], a8[i], isum);
#if defined(__CUDA_ARCH__)
    if (QTraits<TYPE>::HAS_MIN) {
        acc = __fmaf_rn(__fmul_rn(wq.d, as.d), (float) isum, acc);
        return __fmaf_rn(wq.m, as.s, acc);
    } else {
        if (QTraits<TYPE>::OFFSET) isum -= QTraits<TYPE>::OFFSET * (int) as.s;
        return __fmaf_rn(__fmul_rn(wq.d, as.d), (float) isum, acc);
    }
#else
    if (QTraits<TYPE>::HAS_MIN) {
        acc = fmaf(wq.d * as.d, (float) isum, acc);
        return fmaf(wq.m, as.s, acc);
    } else {
        if (QTraits<TYPE>::OFFSET) isum -= QTraits<TYPE>::OFFSET * (int) as.s;
        return fmaf(wq.d * as.d, (float) isum, acc);
    }
#endif
}

}  // namespace rwkv
