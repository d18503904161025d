// Launchers for the non-GEMV kernels of the RWKV eval path: embedding gather + ln0, LayerNorm +
// token shift + mixing, the v6 data-dependent lerp, and the WKV4/5/6/7 recurrences with their
// per-head normalisation and gating fused in. All activations are fp32, column-major [dim, T].
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "gemv.h"   // TraceRec

namespace rwkv {

// x[:, t] = LN(emb[tokens[t]]; ln0)     (rwkv_graph.inc:655-658, 787-790)
cudaError_t launch_embed_ln0(const void * emb, int emb_type, long long emb_pitch, const int * tokens, int T, int C,
                             const float * ln_w, const float * ln_b, float * x, cudaStream_t s);

// LayerNorm + token shift (rwkv_carry_x, rwkv_graph.inc:56-82) + per-arch mixing.
struct LnMixParams {
    TraceRec * trace = nullptr;
    const float * x;          // [C, T] residual stream
    const float * ln_w, * ln_b;
    const float * state_in;   // [C] previous token's LN(x) (att_xx / ffn_xx slot of the input state)
    float * state_out;        // [C] <- LN(x[:, T-1])
    int C, T;
    // formula 0 (v4/v5, rwkv_graph.inc:94-97):  out_j = xx*m_j + (prev - prev*m_j)
    // formula 1 (v6/v7, rwkv_graph.inc:310-311, 406, 520-521, 538):  out_j = (prev - xx)*m_j + xx
    int formula;
    int n_out;
    const float * coef[6];    // [C] each
    float * out[6];           // [C, T] each
    float * out_xx;           // optional [C, T]: LN(x)
    float * out_sx;           // optional [C, T]: prev - xx
    // T == 1, C % 32 == 0 only: out_j additionally as the staged column (act_stage.cuh) of a consumer with weight type q_type[j]
    // (NULL: not wanted); the batch kernel ignores these
    unsigned char * q_out[6];
    int q_type[6];
};
cudaError_t launch_ln_mix(const LnMixParams & p, cudaStream_t s);

// v6 data-dependent lerp (rwkv_graph.inc:323-346): for j in {w,k,v,r,g}
//   m_j[c,t] = sum_i W2[j][c][i] * z[j*mix + i, t];   out_j = (m_j + maa_j[c]) * sx + xx
struct V6LerpParams {
    TraceRec * trace = nullptr;
    const float * w2;         // [5][C][mix] fp32 (ggml ne = [mix, C, 5])
    const float * z;          // [5*mix, T] = tanh(W1 . xxx)
    const float * xx, * sx;   // [C, T]
    const float * maa[5];     // [C]
    float * out[5];           // [C, T]
    int C, T, mix;
    unsigned char * q_out[5]; // T == 1, C % 32 == 0 only: staged columns (act_stage.cuh) for consumers of weight type q_type[j]; all or none
    int q_type[5];
};
cudaError_t launch_v6_lerp(const V6LerpParams & p, cudaStream_t s);

// v4 WKV (rwkv_att_wkv_v4, rwkv_graph.inc:119-161) fused with the r* multiply (:182,195).
struct Wkv4Params {
    TraceRec * trace = nullptr;
    const float * k, * v, * r;        // [C, T]; r already sigmoid-ed
    const float * time_first, * time_decay;   // [C]
    const float * aa_in, * bb_in, * pp_in;    // [C]
    float * aa_out, * bb_out, * pp_out;
    float * y;                        // [C, T] = r * wkv
    int C, T;
    unsigned char * q_out; int q_type;  // T == 1, C % 32 == 0 only: y as a staged column (act_stage.cuh)
};
cudaError_t launch_wkv4(const Wkv4Params & p, cudaStream_t s);

// v5/v6 WKV (ggml_compute_forward_rwkv_wkv6_f32, ggml-cpu.c:11803) + per-head norm + ln_x + gate
// (rwkv_graph.inc:275-289, 370-382).
struct Wkv6Params {
    TraceRec * trace = nullptr;
    const float * r, * k, * v;        // [C, T]
    const float * td;                 // decay: [C, T] if td_per_token else per-channel [C] (v5.2) / per-head [H] (v5.1)
    const float * tf;                 // time_first / time_faaaa: [C] or per-head [H]
    int td_per_token, per_head_scalars;
    const float * state_in;           // [H][S][S]  (i_key major, j_val minor)
    float * state_out;
    const float * lnx_w, * lnx_b;     // [C]
    const float * g;                  // optional [C, T] gate (already silu-ed)
    float * y;                        // [C, T]
    float eps;                        // 1e-5 (v5) or 64e-5 (v6)
    int H, S, T;
    unsigned char * q_out; int q_type;  // T == 1, S % 32 == 0 only: y as a staged column (act_stage.cuh)
};
cudaError_t launch_wkv6(const Wkv6Params & p, cudaStream_t s);

// v7 (rwkv_att_v7, rwkv_graph.inc:432-479 + rwkv_wkv_v7_impl, rwkv_operators_wkv_v7.inc:37-106):
// kk/l2norm, k and v corrections, the recurrence, per-head norm, ln_x, the r*k*r_k bonus and the gate.
struct Wkv7Params {
    TraceRec * trace = nullptr;
    const float * r, * w, * k, * v, * a;   // [C, T]: r raw, w decay, k raw, v raw, a = sigmoid(..)
    const float * g;                  // [C, T] gate
    const float * vgate;              // [C, T] sigmoid(v0 + ...) or NULL for layer 0
    const float * v_first;            // [C, T] layer-0 value (ignored when vgate == NULL)
    float * v_out;                    // optional [C, T]: corrected v (layer 0 stores v_first through it)
    const float * k_k, * k_a, * r_k;  // [C]
    const float * lnx_w, * lnx_b;     // [C]
    const float * state_in;           // [H][S][S]  (i_val major, j_key minor)
    float * state_out;
    float * y;                        // [C, T]
    int H, S, T;
    unsigned char * q_out; int q_type;  // T == 1, S % 32 == 0 only: y as a staged column (act_stage.cuh)
};
cudaError_t launch_wkv7(const Wkv7Params & p, cudaStream_t s);

// Batched multi-sequence decode (batch.cu): the same stages with column t working on the recurrent state of sequence t, which
// lives seq_stride floats after sequence t-1's (p.T = number of sequences; state pointers are sequence 0's).
bool batch_shape_supported(int arch_major, int n_embed, int head_size);
cudaError_t launch_ln_mix_batch(const LnMixParams & p, long long seq_stride, cudaStream_t s);
cudaError_t launch_wkv6_batch(const Wkv6Params & p, long long seq_stride, cudaStream_t s);
cudaError_t launch_wkv4_batch(const Wkv4Params & p, long long seq_stride, cudaStream_t s);
cudaError_t launch_wkv7_batch(const Wkv7Params & p, long long seq_stride, cudaStream_t s);

// Layer-pipeline hand-off over peer memory (pipe.cu). A PipeBox sits at the start of every stage's mailbox allocation, followed by
// PIPE_SLOTS data areas of slot_floats floats each.
constexpr int PIPE_SLOTS = 2;
constexpr int PIPE_CTAS = 8;               // every hand-off launch has this many CTAs (the device-side item counters rely on it)
struct PipeBox {
    unsigned long long credit;             // written by the NEXT stage: items of mine it has drained
    unsigned long long pad0[15];
    unsigned long long full[PIPE_SLOTS];   // written by the PREVIOUS stage: item + 1 that the slot holds
    unsigned long long pad1[16 - PIPE_SLOTS];
};
cudaError_t launch_pipe_recv(PipeBox * mine, PipeBox * prev, unsigned long long * counters, size_t slot_floats, float * x, size_t nx, float * v, size_t nv, cudaStream_t s);
cudaError_t launch_pipe_send(PipeBox * mine, PipeBox * next, unsigned long long * counters, size_t slot_floats, float * x, size_t nx, float * v, size_t nv, cudaStream_t s);

// On-device sampling from the logits of the last evaluated token (reference python/sampling.py:10-52); see sampling.cu.
struct SampleParams {
    const float * logits;             // [n_vocab] device
    int n_vocab;
    float temperature, top_p;
    double u;                         // uniform number in [0, 1) drawn by the caller
    const uint32_t * bias_ids;        // device, n_bias entries (logit_bias of the reference), or NULL
    const float * bias_values;
    int n_bias;
    float * scratch;                  // [n_vocab] device, used only when n_bias > 0
    uint32_t * token_out;             // device
    float * prob_out;                 // optional device: probability the chosen token had
};
cudaError_t launch_sample(const SampleParams & p, cudaStream_t s);

}  // namespace rwkv
