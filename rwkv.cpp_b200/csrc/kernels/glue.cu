// Embedding gather + ln0, LayerNorm + token shift + mixing, and the v6 data-dependent lerp.
// These are the small fp32 element-wise stages between the GEMVs; each is one launch over [C, T].
#include "ops.h"
#include "gemv.h"          // g_kernel_launches
#include "../formats.h"

#include <cuda_fp16.h>

namespace rwkv {
namespace {

constexpr int GLUE_THREADS = 256;
constexpr int GLUE_WARPS = GLUE_THREADS / 32;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ double block_sum_d(double v, double * scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum_d(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0;
        for (int i = 0; i < GLUE_WARPS; i++) s += scratch[i];
        scratch[GLUE_WARPS] = s;
    }
    __syncthreads();
    return scratch[GLUE_WARPS];
}

// dst[c] = LN(src)[c] * w[c] + b[c]; statistics as ggml_compute_forward_norm_f32 (ggml-cpu.c:6906-6925),
// weight/bias as rwkv_layer_norm (rwkv_operators.inc:93-97). src may be shared or global.
__device__ void layer_norm_to(const float * src, const float * w, const float * b, int C, float * dst, double * scratch) {
    double s = 0;
    for (int c = threadIdx.x; c < C; c += GLUE_THREADS) s += (double) src[c];
    const float mean = (float) (block_sum_d(s, scratch) / C);
    double s2 = 0;
    for (int c = threadIdx.x; c < C; c += GLUE_THREADS) { float v = src[c] - mean; s2 += (double) (v * v); }
    const float var = (float) (block_sum_d(s2, scratch) / C);
    const float scale = 1.0f / sqrtf(var + 1e-5f);
    for (int c = threadIdx.x; c < C; c += GLUE_THREADS)
        dst[c] = __fadd_rn(__fmul_rn(__fmul_rn(src[c] - mean, scale), w[c]), b[c]);
    __syncthreads();
}

__global__ void __launch_bounds__(GLUE_THREADS) embed_ln0_kernel(const uint8_t * emb, int emb_type, long long pitch, const int * tokens, int C,
                                                                  const float * ln_w, const float * ln_b, float * x) {
    extern __shared__ float sh[];
    __shared__ double scratch[GLUE_WARPS + 1];
    const int t = blockIdx.x;
    const uint8_t * row = emb + (size_t) tokens[t] * (size_t) pitch;
    for (int c = threadIdx.x; c < C; c += GLUE_THREADS)
        sh[c] = (emb_type == DT_F16) ? __half2float(reinterpret_cast<const __half *>(row)[c]) : reinterpret_cast<const float *>(row)[c];
    __syncthreads();
    layer_norm_to(sh, ln_w, ln_b, C, x + (size_t) t * C, scratch);
}

__global__ void __launch_bounds__(GLUE_THREADS) ln_mix_kernel(const LnMixParams p) {
    extern __shared__ float sh[];
    __shared__ double scratch[GLUE_WARPS + 1];
    const int C = p.C, t = blockIdx.x;
    float * xx = sh;            // LN(x[:, t])
    float * prev = sh + C;      // LN(x[:, t-1]) or the carried state
    layer_norm_to(p.x + (size_t) t * C, p.ln_w, p.ln_b, C, xx, scratch);
    if (t == 0) {
        for (int c = threadIdx.x; c < C; c += GLUE_THREADS) prev[c] = p.state_in[c];
        __syncthreads();
    } else {
        layer_norm_to(p.x + (size_t) (t - 1) * C, p.ln_w, p.ln_b, C, prev, scratch);
    }
    for (int c = threadIdx.x; c < C; c += GLUE_THREADS) {
        const float a = xx[c], b = prev[c];
        const size_t o = (size_t) t * C + c;
        if (p.formula == 0) {
#pragma unroll
            for (int j = 0; j < 6; j++)
                if (j < p.n_out) {
                    const float m = p.coef[j][c];
                    p.out[j][o] = __fadd_rn(__fmul_rn(a, m), __fsub_rn(b, __fmul_rn(b, m)));
                }
        } else {
            const float sx = __fsub_rn(b, a);
#pragma unroll
            for (int j = 0; j < 6; j++)
                if (j < p.n_out) p.out[j][o] = __fadd_rn(__fmul_rn(sx, p.coef[j][c]), a);
            if (p.out_sx) p.out_sx[o] = sx;
        }
        if (p.out_xx) p.out_xx[o] = a;
        if (t == p.T - 1) p.state_out[c] = a;
    }
}

// 8 lanes per channel, 32 channels per CTA: each (j, channel) row of W2 is `mix` contiguous floats.
__global__ void __launch_bounds__(GLUE_THREADS) v6_lerp_kernel(const V6LerpParams p) {
    extern __shared__ float zs[];   // [5*mix]
    const int t = blockIdx.y, mix = p.mix, C = p.C;
    for (int i = threadIdx.x; i < 5 * mix; i += GLUE_THREADS) zs[i] = p.z[(size_t) t * 5 * mix + i];
    __syncthreads();
    const int grp = threadIdx.x >> 3, sub = threadIdx.x & 7;
    const int c = blockIdx.x * 32 + grp;
    const bool live = c < C;
    const size_t o = (size_t) t * C + (live ? c : 0);
    const float sx = live ? p.sx[o] : 0.f, xx = live ? p.xx[o] : 0.f;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float acc = 0.f;
        if (live) {
            const float * wrow = p.w2 + ((size_t) j * C + c) * mix;
            const float * zj = zs + j * mix;
            if ((mix & 3) == 0) {
                for (int i4 = sub; i4 < mix / 4; i4 += 8) {
                    float4 w = __ldg(reinterpret_cast<const float4 *>(wrow) + i4);
                    float4 z = reinterpret_cast<const float4 *>(zj)[i4];
                    acc = __fmaf_rn(w.x, z.x, acc); acc = __fmaf_rn(w.y, z.y, acc);
                    acc = __fmaf_rn(w.z, z.z, acc); acc = __fmaf_rn(w.w, z.w, acc);
                }
            } else {
                for (int i = sub; i < mix; i += 8) acc = __fmaf_rn(wrow[i], zj[i], acc);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (live && sub == 0) p.out[j][o] = __fadd_rn(__fmul_rn(__fadd_rn(acc, p.maa[j][c]), sx), xx);
    }
}

}  // namespace

cudaError_t launch_embed_ln0(const void * emb, int emb_type, long long emb_pitch, const int * tokens, int T, int C,
                             const float * ln_w, const float * ln_b, float * x, cudaStream_t s) {
    embed_ln0_kernel<<<T, GLUE_THREADS, (size_t) C * sizeof(float), s>>>(reinterpret_cast<const uint8_t *>(emb), emb_type, emb_pitch, tokens, C, ln_w, ln_b, x);
    g_kernel_launches++;
    return cudaGetLastError();
}

cudaError_t launch_ln_mix(const LnMixParams & p, cudaStream_t s) {
    ln_mix_kernel<<<p.T, GLUE_THREADS, (size_t) 2 * p.C * sizeof(float), s>>>(p);
    g_kernel_launches++;
    return cudaGetLastError();
}

cudaError_t launch_v6_lerp(const V6LerpParams & p, cudaStream_t s) {
    dim3 grid((p.C + 31) / 32, p.T);
    v6_lerp_kernel<<<grid, GLUE_THREADS, (size_t) 5 * p.mix * sizeof(float), s>>>(p);
    g_kernel_launches++;
    return cudaGetLastError();
}

}  // namespace rwkv
