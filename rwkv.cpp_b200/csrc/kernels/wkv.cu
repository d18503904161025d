// WKV time-mixing recurrences for RWKV v4 / v5 / v6 / v7, each fused with the normalisation and
// gating that follows it in the reference graph. The recurrent state stays in registers for the
// whole chunk; one CTA per head (v5+), one thread per state column (v5/v6) or row (v7).
#include "ops.h"
#include "gemv.h"   // g_kernel_launches

namespace rwkv {
namespace {

// ---------------------------------------------------------------------------------------------
// v4: rwkv_att_wkv_v4 (rwkv_graph.inc:119-161). One thread per channel, sequential over tokens.
// Every ggml node is its own rounding, so no FMA contraction here.
// ---------------------------------------------------------------------------------------------
__global__ void wkv4_kernel(const Wkv4Params p) {
    trace_begin(p.trace);
    pdl_prologue();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= p.C) return;
    float aa = p.aa_in[c], bb = p.bb_in[c], pp = p.pp_in[c];
    const float tf = p.time_first[c], td = p.time_decay[c];
    for (int t = 0; t < p.T; t++) {
        const size_t o = (size_t) t * p.C + c;
        const float k = p.k[o], v = p.v[o];
        float ww = __fadd_rn(tf, k);
        float qq = fmaxf(pp, ww);
        float e1 = expf(__fsub_rn(pp, qq)), e2 = expf(__fsub_rn(ww, qq));
        const float a = __fadd_rn(__fmul_rn(e1, aa), __fmul_rn(e2, v));
        const float b = __fadd_rn(__fmul_rn(e1, bb), e2);
        ww = __fadd_rn(pp, td);
        qq = fmaxf(ww, k);
        e1 = expf(__fsub_rn(ww, qq));
        e2 = expf(__fsub_rn(k, qq));
        aa = __fadd_rn(__fmul_rn(e1, aa), __fmul_rn(e2, v));
        bb = __fadd_rn(__fmul_rn(e1, bb), e2);
        pp = qq;
        p.y[o] = __fmul_rn(p.r[o], __fdiv_rn(a, b));
    }
    p.aa_out[c] = aa; p.bb_out[c] = bb; p.pp_out[c] = pp;
    trace_end(p.trace);
}

// Per-head normalisation of S values held one per thread: ggml_norm over a head
// (ggml-cpu.c:6906-6925: mean and variance summed sequentially in double).
template <int S>
__device__ __forceinline__ float head_norm(float y, float eps, float * red) {
    const int j = threadIdx.x;
    __syncthreads();
    red[j] = y;
    __syncthreads();
    double s = 0;
#pragma unroll 8
    for (int i = 0; i < S; i++) s += (double) red[i];
    const float mean = (float) (s / S);
    const float v = y - mean;
    __syncthreads();
    red[j] = v * v;
    __syncthreads();
    double s2 = 0;
#pragma unroll 8
    for (int i = 0; i < S; i++) s2 += (double) red[i];
    const float var = (float) (s2 / S);
    return v * (1.0f / sqrtf(var + eps));
}

// ---------------------------------------------------------------------------------------------
// v5 / v6: ggml_compute_forward_rwkv_wkv6_f32 (ggml-cpu.c:11870-11905). Thread j owns state column
// S[.][j]; per token  kv = v_j*k_i;  y_j += (kv*tf_i + S_ij) * r_i;  S_ij = S_ij*td_i + kv
// with the same FMA grouping as the reference's vector path.
// ---------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(S) wkv6_kernel(const Wkv6Params p) {
    __shared__ float sk[S], sr[S], sd[S], sf[S], red[S];
    trace_begin(p.trace);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int h = blockIdx.x, j = threadIdx.x, C = p.H * S;
    // the recurrent state (written by the previous token's pass) and the parameters do not depend on the previous
    // kernel: pull them in before the programmatic-dependency wait
    float st[S];
#pragma unroll
    for (int i = 0; i < S; i++) st[i] = p.state_in[((size_t) h * S + i) * S + j];
    sf[j] = p.per_head_scalars ? p.tf[h] : p.tf[h * S + j];
    if (!p.td_per_token) sd[j] = p.per_head_scalars ? p.td[h] : p.td[h * S + j];
    const float lw = p.lnx_w[h * S + j], lb = p.lnx_b[h * S + j];
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int t = 0; t < p.T; t++) {
        const size_t o = (size_t) t * C + h * S + j;
        __syncthreads();
        sk[j] = p.k[o];
        sr[j] = p.r[o];
        if (p.td_per_token) sd[j] = p.td[o];
        __syncthreads();
        const float vj = p.v[o];
        float y = 0.f;
#pragma unroll
        for (int i = 0; i < S; i++) {
            const float kv = __fmul_rn(vj, sk[i]);
            const float temp = __fmaf_rn(kv, sf[i], st[i]);
            y = __fmaf_rn(temp, sr[i], y);
            st[i] = __fmaf_rn(st[i], sd[i], kv);
        }
        float n = head_norm<S>(y, p.eps, red);
        n = __fadd_rn(__fmul_rn(n, lw), lb);
        if (p.g) n = __fmul_rn(n, p.g[o]);
        p.y[o] = n;
    }
#pragma unroll
    for (int i = 0; i < S; i++) p.state_out[((size_t) h * S + i) * S + j] = st[i];
    trace_end(p.trace);
}

// ---------------------------------------------------------------------------------------------
// v7: rwkv_att_v7 (rwkv_graph.inc:432-479) around rwkv_wkv_v7_impl (rwkv_operators_wkv_v7.inc:63-101).
// Thread i owns state row S[i][.] (value index i, key index j).
// ---------------------------------------------------------------------------------------------
template <int S>
__global__ void __launch_bounds__(S) wkv7_kernel(const Wkv7Params p) {
    __shared__ float sr[S], sw[S], sk[S], sa[S], sb[S], red[S];
    trace_begin(p.trace);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int h = blockIdx.x, i = threadIdx.x, C = p.H * S, c = h * S + i;
    float st[S];
#pragma unroll
    for (int j = 0; j < S; j++) st[j] = p.state_in[((size_t) h * S + i) * S + j];
    const float kk_w = p.k_k[c], ka_w = p.k_a[c], rk_w = p.r_k[c];
    const float lw = p.lnx_w[c], lb = p.lnx_b[c];
    asm volatile("griddepcontrol.wait;" ::: "memory");
    for (int t = 0; t < p.T; t++) {
        const size_t o = (size_t) t * C + c;
        const float r = p.r[o], w = p.w[o], k0 = p.k[o], a = p.a[o];
        float v = p.v[o];
        // kk = l2norm_head(k * k_k)  (rwkv_graph.inc:433-434, rwkv_operators.inc:61-76: float sum, eps 1e-12)
        const float kk0 = __fmul_rn(k0, kk_w);
        __syncthreads();
        red[i] = kk0;
        __syncthreads();
        float ss = 0.f;
#pragma unroll 8
        for (int j = 0; j < S; j++) ss = __fmaf_rn(red[j], red[j], ss);
        const float kk = __fmul_rn(kk0, 1.0f / fmaxf(sqrtf(ss), 1e-12f));
        // k = k + (a*ka - ka), ka = k*k_a   (rwkv_graph.inc:436-437)
        const float ka = __fmul_rn(k0, ka_w);
        const float k = __fadd_rn(k0, __fsub_rn(__fmul_rn(a, ka), ka));
        // v = v + (v_first - v) * gate   (rwkv_graph.inc:440-453)
        if (p.vgate) v = __fadd_rn(v, __fmul_rn(__fsub_rn(p.v_first[o], v), p.vgate[o]));
        if (p.v_out) p.v_out[o] = v;
        __syncthreads();
        sr[i] = r; sw[i] = w; sk[i] = k; sa[i] = -kk; sb[i] = __fmul_rn(kk, a);
        __syncthreads();
        float dot_a = 0.f;
#pragma unroll
        for (int j = 0; j < S; j++) dot_a = __fmaf_rn(sa[j], st[j], dot_a);
        float y = 0.f;
#pragma unroll
        for (int j = 0; j < S; j++) {
            const float kv = __fmul_rn(v, sk[j]);
            st[j] = __fmaf_rn(dot_a, sb[j], __fmaf_rn(st[j], sw[j], kv));
            y = __fmaf_rn(st[j], sr[j], y);
        }
        float n = head_norm<S>(y, 64e-5f, red);
        n = __fadd_rn(__fmul_rn(n, lw), lb);
        // + v * sum_head(k*r*r_k)   (rwkv_graph.inc:472-477; ggml_sum_rows accumulates in double)
        __syncthreads();
        red[i] = __fmul_rn(__fmul_rn(k, r), rk_w);
        __syncthreads();
        double rk = 0;
#pragma unroll 8
        for (int j = 0; j < S; j++) rk += (double) red[j];
        n = __fadd_rn(n, __fmul_rn(v, (float) rk));
        p.y[o] = __fmul_rn(n, p.g[o]);
    }
#pragma unroll
    for (int j = 0; j < S; j++) p.state_out[((size_t) h * S + i) * S + j] = st[j];
    trace_end(p.trace);
}

}  // namespace

cudaError_t launch_wkv4(const Wkv4Params & p_in, cudaStream_t s) {
    Wkv4Params p = p_in;
    p.trace = trace_slot("wkv4");
    const int threads = 128;
    g_kernel_launches++;
    return launch_pdl(wkv4_kernel, dim3((p.C + threads - 1) / threads), dim3(threads), 0, s, p);
}

#define RWKV_DISPATCH_HEAD_SIZE(S_, KERNEL, PARAMS, STREAM)                              \
    switch (S_) {                                                                        \
        case 8: return launch_pdl(KERNEL<8>, dim3(PARAMS.H), dim3(8), 0, STREAM, PARAMS);         \
        case 16: return launch_pdl(KERNEL<16>, dim3(PARAMS.H), dim3(16), 0, STREAM, PARAMS);      \
        case 32: return launch_pdl(KERNEL<32>, dim3(PARAMS.H), dim3(32), 0, STREAM, PARAMS);      \
        case 64: return launch_pdl(KERNEL<64>, dim3(PARAMS.H), dim3(64), 0, STREAM, PARAMS);      \
        case 128: return launch_pdl(KERNEL<128>, dim3(PARAMS.H), dim3(128), 0, STREAM, PARAMS);   \
        default: return cudaErrorInvalidValue;                                           \
    }

cudaError_t launch_wkv6(const Wkv6Params & p_in, cudaStream_t s) {
    Wkv6Params p = p_in;
    p.trace = trace_slot("wkv6");
    g_kernel_launches++;
    RWKV_DISPATCH_HEAD_SIZE(p.S, wkv6_kernel, p, s)
}

cudaError_t launch_wkv7(const Wkv7Params & p_in, cudaStream_t s) {
    Wkv7Params p = p_in;
    p.trace = trace_slot("wkv7");
    g_kernel_launches++;
    RWKV_DISPATCH_HEAD_SIZE(p.S, wkv7_kernel, p, s)
}

}  // namespace rwkv
