// Forward pass over a resident model: the per-architecture launch sequences.
// Math follows rwkv_graph.inc (serial graph :611-720, sequence graph :744-866); every stage below
// names the graph lines it covers. All launches go to ctx->stream; nothing here blocks the host
// except the token upload.
#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "kernels/act_stage.cuh"
#include "kernels/gemv.h"
#include "kernels/ops.h"

namespace rwkv {

namespace {

// rwkv_eval with host state buffers: pipeline the state copies against layer groups? (RWKV_B200_OVERLAP=0/1, rwkv_b200_set_overlap)
constexpr bool OVERLAP_DEFAULT = true;

struct Scratch {      // carve-up of ctx->scratch for T tokens
    float * x, * xx, * sx;
    float * mix[6];
    float * r, * k, * v, * g, * w, * a, * vgate, * v_first, * y;
    float * ffn_k, * ffn_r;
    float * lora[4];
};

constexpr int BATCH_TC_MIN = 16;
// Smallest batch (in weights) that goes to the tensor cores in a pass of >= 32 tokens: 0 = every eligible batch. A tensor-core launch
// has a fixed cost of ~40 us (43 us for the 160 x 4096 LoRA matrix of a v6 layer), but the multi-column GEMV is worse for such a
// matrix at T = 128 (200 us: 32 column groups): measured 20.3 ms per 7B chunk with a 2^20 threshold against 15.0 ms with none
// (profiles/r2_c8_pf_*.json). RWKV_B200_TC_MIN_WEIGHTS overrides (A/B aid).
long long tc_min_weights() {
    static const long long v = [] { const char * e = getenv("RWKV_B200_TC_MIN_WEIGHTS"); return e ? atoll(e) : 0ll; }();
    return v;
}

struct Dims { size_t C, F, R; };

Dims model_dims(const Model & m) {
    Dims d{(size_t) m.n_embed, 0, 0};
    for (int i = m.layer_begin; i < m.layer_end; i++) {
        const Layer & L = m.layers[i];
        if ((size_t) L.ffn_key.M > d.F) d.F = L.ffn_key.M;
        const int ranks[] = {L.att_maa_w1.M, L.att_decay_w1.M, L.att_w1.M, L.att_a1.M, L.att_g1.M, L.att_v1.M};
        for (int r : ranks) if ((size_t) r > d.R) d.R = r;
    }
    if (d.R == 0) d.R = 1;
    return d;
}

// Layers [l0, l1) of segment `seg` (0 = every resident layer).
void segment_range(const Context * ctx, int seg, int & l0, int & l1) {
    const Model & m = *ctx->model;
    l0 = m.layer_begin; l1 = m.layer_end;
    if (seg <= 0) return;
    const int n = m.layer_end - m.layer_begin, G = ctx->n_segments;
    // Graded groups: 1, 2, 3, 5, 10, 6, 3, 2 thirty-seconds of the layers. The upload of the first group's state slice and the download
    // of the last group's are the two copies nothing can overlap, so those groups are small; a group may grow by about the ratio of
    // compute time to copy time per layer (1.7 at 7B: 75 us vs 43 us) over its predecessor without starving the kernels, and shrink
    // the same way towards the end without queueing the downloads. A pipeline model with those two rates puts a 7B token at 2.60 ms
    // against 2.81 ms for eight even groups (2.46 ms: kernels alone). RWKV_B200_EVEN_SEGMENTS=1: the even split of round 1 (A/B aid).
    static const bool even = getenv("RWKV_B200_EVEN_SEGMENTS") != nullptr;
    static const int cum32[9] = {0, 1, 3, 6, 11, 21, 27, 30, 32};
    if (even || G != 8 || n < 16) {
        l0 = m.layer_begin + (int) ((long long) n * (seg - 1) / G);
        l1 = m.layer_begin + (int) ((long long) n * seg / G);
        return;
    }
    auto cut = [&](int s) -> int { return (int) (((long long) cum32[s] * n + 16) / 32); };      // strictly increasing for n >= 16
    l0 = m.layer_begin + cut(seg - 1);
    l1 = m.layer_begin + cut(seg);
}

size_t scratch_floats_for(const Model & m, int T) {
    Dims d = model_dims(m);
    return ((size_t) 19 * d.C + d.F + 4 * d.R) * (size_t) T + 64;
}

Scratch carve(const Model & m, float * base, int T) {
    Dims d = model_dims(m);
    Scratch s;
    float * p = base;
    auto take = [&](size_t n) { float * q = p; p += n * (size_t) T; return q; };
    s.x = take(d.C); s.xx = take(d.C); s.sx = take(d.C);
    for (int i = 0; i < 6; i++) s.mix[i] = take(d.C);
    s.r = take(d.C); s.k = take(d.C); s.v = take(d.C); s.g = take(d.C); s.w = take(d.C); s.a = take(d.C);
    s.vgate = take(d.C); s.v_first = take(d.C); s.y = take(d.C);
    s.ffn_r = take(d.C);
    s.ffn_k = take(d.F);
    for (int i = 0; i < 4; i++) s.lora[i] = take(d.R);
    return s;
}

// ---- GEMV batch builder -------------------------------------------------------------------------
struct Batch {
    GemvBatch b;
    explicit Batch(int T) { memset(&b, 0, sizeof(b)); b.T = T; }
    GemvProblem & add(const DevMatrix & W, const float * x, float * y, int epi = EPI_NONE) {
        GemvProblem & p = b.p[b.n++];
        p.W = W.data; p.Wt = W.tiled; p.pitch = W.pitch; p.type = W.type; p.K = W.K; p.M = W.M;
        p.x = x; p.ldx = W.K;
        p.y = y; p.ldy = W.M;
        p.epi = epi; p.pro = PRO_NONE;
        return p;
    }
};

// ---- staged activation columns (act_stage.cuh) --------------------------------------------------------------------------------------
// Single-token passes: the kernel that produces a GEMV's input vector (LayerNorm + mix, the v6 lerp, a WKV kernel) also writes
// it in the layout the streaming GEMV keeps in shared memory, so the consumer copies ~5 KB instead of quantising 16 KB of fp32 on
// its critical path. One slot per (producer, output); a slot serves every consumer whose weight type multiplies the same staged format.
constexpr int XQ_SLOTS = 16;
enum XqSlot { XQ_ATT_MIX = 0 /* .. 5 */, XQ_LERP = 6 /* .. 10 */, XQ_WKV = 11, XQ_FFN_MIX = 12 /* .. 13 */ };

bool handoff_enabled() {
    static const bool on = getenv("RWKV_B200_NO_XQ") == nullptr && getenv("RWKV_B200_GENERIC_GEMV") == nullptr;
    return on;
}
// The slot to fill for a consumer matrix W of this pass, or NULL when the pass / shape cannot use the hand-off.
unsigned char * xq_slot(const Context * ctx, int T, int slot, const DevMatrix & W) {
    if (!handoff_enabled() || T != 1 || ctx->batch_stride || !ctx->xq || !W.data || W.K % 32 != 0 || W.type == DT_F32 || act::stage_class(W.type) == act::SC_NONE) return nullptr;
    if (act::bytes_per_column(W.type, W.K) > ctx->xq_slot_bytes) return nullptr;
    return ctx->xq + (size_t) slot * ctx->xq_slot_bytes;
}
// Point problem p at a staged copy of its input that was written for weight type q_type.
void xq_bind(GemvProblem & p, const unsigned char * q, int q_type) {
    if (q && act::stage_class(p.type) == act::stage_class(q_type)) p.xq = q;
}
#define CUDA_OK(ctx, call)                                                                               \
    do { cudaError_t _e = (call);                                                                        \
         RWKV_CHECK((ctx)->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, _e == cudaSuccess,       \
                    "CUDA error: %s", cudaGetErrorString(_e)); } while (0)

// Passes of >= 32 tokens send every problem the tensor-core kernel can take (any weight type but F32, K % 64 == 0) to
// gemm_tc.cu; the rest (and all shorter passes) use the batch-invariant GEMV, so serial == sequence stays bit-exact
// wherever the reference's tests compare states.
bool launch_batch(Context * ctx, GemvBatch & b) {
    // tensor cores: chunks of >= 32 tokens; batch contexts (one token of each of B sequences: nothing to stay bit-identical with
    // across pass sizes) from 16 sequences on, where the multi-column SIMT consumer is instruction-bound
    if ((b.T >= 32 || (ctx->batch_n > 0 && b.T >= BATCH_TC_MIN)) && ctx->use_tensor_cores && ctx->act16) {
        GemvBatch tcb, rest;
        memset(&tcb, 0, sizeof(tcb)); memset(&rest, 0, sizeof(rest));
        tcb.T = rest.T = b.T;
        long long tc_weights = 0;
        for (int i = 0; i < b.n; i++) {
            if (gemm_tc_supported(b.p[i], b.T)) { tcb.p[tcb.n++] = b.p[i]; tc_weights += (long long) b.p[i].M * b.p[i].K; }
            else rest.p[rest.n++] = b.p[i];
        }
        if (tc_weights < tc_min_weights()) {
            for (int i = 0; i < tcb.n; i++) rest.p[rest.n++] = tcb.p[i];
            tcb.n = 0;
        }
        if (tcb.n) CUDA_OK(ctx, gemm_tc_launch(tcb, ctx->model->dev, ctx->stream, ctx->act16, ctx->act16_bytes));
        if (rest.n) CUDA_OK(ctx, gemv_launch(rest, ctx->model->dev, ctx->stream));
        return true;
    }
    CUDA_OK(ctx, gemv_launch(b, ctx->model->dev, ctx->stream));
    return true;
}

bool run_batch(Context * ctx, Batch & batch) {
    if (!ctx->profiling) return launch_batch(ctx, batch.b);
    Context::ProfRecord rec{};
    CUDA_OK(ctx, cudaEventCreate(&rec.start));
    CUDA_OK(ctx, cudaEventCreate(&rec.stop));
    for (int i = 0; i < batch.b.n; i++)
        rec.bytes += (double) tensor_nbytes(batch.b.p[i].type, (uint64_t) batch.b.p[i].K, (uint64_t) batch.b.p[i].M, 1);
    CUDA_OK(ctx, cudaEventRecord(rec.start, ctx->stream));
    if (!launch_batch(ctx, batch.b)) return false;
    CUDA_OK(ctx, cudaEventRecord(rec.stop, ctx->stream));
    ctx->prof.push_back(rec);
    return true;
}

// The three stages that touch the recurrent state: a batch context (one token of each of batch_n sequences) uses the kernels of
// batch.cu, where column t works on sequence t's state; everything else is identical for both kinds of context.
cudaError_t do_ln_mix(Context * ctx, const LnMixParams & lp) {
    return ctx->batch_stride ? launch_ln_mix_batch(lp, ctx->batch_stride, ctx->stream) : launch_ln_mix(lp, ctx->stream);
}
cudaError_t do_wkv6(Context * ctx, const Wkv6Params & wp) {
    return ctx->batch_stride ? launch_wkv6_batch(wp, ctx->batch_stride, ctx->stream) : launch_wkv6(wp, ctx->stream);
}
cudaError_t do_wkv4(Context * ctx, const Wkv4Params & wp) {
    return ctx->batch_stride ? launch_wkv4_batch(wp, ctx->batch_stride, ctx->stream) : launch_wkv4(wp, ctx->stream);
}

// LayerNorm + token shift + mix in front of a batch of GEMVs: its own launch (ln_mix_kernel); *staged = the mixed vectors also left
// as staged columns (never, today). Three variations were built, measured at 7B Q5_1 and removed in round 2:
// (a) fused into the CONSUMING GEMV, every CTA recomputing the LayerNorm in its prologue: 4.24 vs 3.14 ms per token
//     (profiles/r2_trace_decode_c7_fuseln.log);
// (b) as the tail job of the PRODUCING GEMV (att.output / ffn.value), run by the CTA that draws the last ticket: 3.62 vs 2.86 ms
//     (profiles/r2_c11_ab_*.json) -- every CTA pays a device-wide fence + ticket before it may retire, and the one CTA left behind does
//     the LayerNorm of 4096 channels with 256 threads where this kernel uses 1024 and has its parameter loads in flight before the
//     dependency wait;
// (c) ln_mix_kernel emitting the staged columns of its consumers: 2.95 vs 2.86 ms (profiles/r2_c12_ab_default.json).
// And for passes of >= 32 tokens: (d) ln_mix_kernel writing the fp16 GEMM operands of its consumers itself, so that two of the six
//     convert_f16 launches of a v6 layer disappear: 10.74 vs 10.85 ms per 7B chunk (profiles/r2_c21_pf_ln16.json / _noln16.json), and
//     the tail of the state went NaN in the shift-state-3e5 robustness test of the 64-channel fixture
//     (profiles/r2_c21_parity_ln16_nan_at_3e5.log; not understood). 1 % was not worth an unexplained failure: not kept.
bool ln_mix_then(Context * ctx, const LnMixParams & lp, bool * staged) {
    *staged = false;
    CUDA_OK(ctx, do_ln_mix(ctx, lp));
    return true;
}

// LayerNorm + mix parameters of the channel-mixing block (rwkv_ffn_v4_v5 :484-511, rwkv_ffn_v6 :513-531, rwkv_ffn_v7 :533-543).
LnMixParams ffn_ln(const Context * ctx, const Layer & L, const Scratch & s, int T, const float * st_in, float * st_out) {
    const Model & m = *ctx->model;
    LnMixParams lp{};
    lp.x = s.x; lp.ln_w = L.ln2_w.data; lp.ln_b = L.ln2_b.data;
    lp.state_in = st_in; lp.state_out = st_out; lp.C = m.n_embed; lp.T = T;
    if (m.arch_major == 7) {
        lp.formula = 1; lp.n_out = 1; lp.coef[0] = L.ffn_x_k.data; lp.out[0] = s.mix[0];
    } else if (m.arch_major == 6) {
        lp.formula = 1; lp.n_out = 2;
        lp.coef[0] = L.ffn_maa_k.data; lp.out[0] = s.mix[0];
        lp.coef[1] = L.ffn_maa_r.data; lp.out[1] = s.mix[1];
    } else {
        lp.formula = 0; lp.n_out = 2;
        lp.coef[0] = L.ffn_time_mix_k.data; lp.out[0] = s.mix[0];
        lp.coef[1] = L.ffn_time_mix_r.data; lp.out[1] = s.mix[1];
    }
    lp.q_out[0] = xq_slot(ctx, T, XQ_FFN_MIX, L.ffn_key); lp.q_type[0] = L.ffn_key.type;
    if (m.arch_major != 7) { lp.q_out[1] = xq_slot(ctx, T, XQ_FFN_MIX + 1, L.ffn_receptance); lp.q_type[1] = L.ffn_receptance.type; }
    return lp;
}

// Channel mixing, all versions.
bool ffn(Context * ctx, const Layer & L, const Scratch & s, int T, const LnMixParams & lp) {
    const Model & m = *ctx->model;
    const int C = m.n_embed;
    {
        Batch b(T);
        GemvProblem & pk = b.add(L.ffn_key, s.mix[0], s.ffn_k, EPI_RELU_SQR);
        GemvProblem * pr = (m.arch_major != 7) ? &b.add(L.ffn_receptance, s.mix[1], s.ffn_r, EPI_SIGMOID) : nullptr;
        bool staged = false;
        if (!ln_mix_then(ctx, lp, &staged)) return false;
        if (staged) { xq_bind(pk, lp.q_out[0], lp.q_type[0]); if (pr) xq_bind(*pr, lp.q_out[1], lp.q_type[1]); }
        if (!run_batch(ctx, b)) return false;
    }
    {
        Batch b(T);
        GemvProblem & p = b.add(L.ffn_value, s.ffn_k, s.x, m.arch_major == 7 ? EPI_ADD : EPI_MUL_ADD);
        p.res = s.x; p.ldres = C;
        p.gate = s.ffn_r; p.ldgate = C;
        if (!run_batch(ctx, b)) return false;
    }
    return true;
}

// yq: y as a staged column (written by the WKV kernel) or NULL
bool att_output(Context * ctx, const Layer & L, const Scratch & s, int T, const unsigned char * yq) {
    Batch b(T);
    GemvProblem & p = b.add(L.att_output, s.y, s.x, EPI_ADD);   // x + Wo.y  (:182/:291/:384/:481 + residual :667-679)
    p.res = s.x; p.ldres = ctx->model->n_embed;
    xq_bind(p, yq, L.att_output.type);
    return run_batch(ctx, b);
}

// LayerNorm + mix parameters of the time-mixing block of layer `layer`, per architecture (:94-97, :306-311, :400-413), with the staged
// columns of the consumers of each mixed vector.
LnMixParams att_ln(const Context * ctx, const Layer & L, int layer, const Scratch & s, int T, const float * st_in, float * st_out) {
    const Model & m = *ctx->model;
    const int C = m.n_embed;
    LnMixParams lp{};
    lp.x = s.x; lp.ln_w = L.ln1_w.data; lp.ln_b = L.ln1_b.data;
    lp.state_in = st_in + C; lp.state_out = st_out + C; lp.C = C; lp.T = T;
    const DevMatrix * cons[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};      // first consumer of out[j]
    if (m.arch_major == 7) {
        lp.formula = 1; lp.n_out = 6;
        for (int j = 0; j < 6; j++) { lp.coef[j] = L.att_x_rwkvag.data + (size_t) j * C; lp.out[j] = s.mix[j]; }   // r w k v a g
        cons[0] = &L.att_receptance; cons[1] = &L.att_w1; cons[2] = &L.att_key; cons[3] = &L.att_value; cons[4] = &L.att_a1; cons[5] = &L.att_g1;
    } else if (m.arch_major == 6) {
        lp.formula = 1; lp.n_out = 1; lp.coef[0] = L.att_maa_x.data; lp.out[0] = s.mix[0];
        lp.out_xx = s.xx; lp.out_sx = s.sx;
        cons[0] = &L.att_maa_w1;
    } else {
        const bool v52 = m.arch_major == 5 && m.arch_minor >= 2;
        lp.formula = 0; lp.n_out = v52 ? 4 : 3;
        lp.coef[0] = L.att_time_mix_k.data; lp.out[0] = s.mix[0];
        lp.coef[1] = L.att_time_mix_v.data; lp.out[1] = s.mix[1];
        lp.coef[2] = L.att_time_mix_r.data; lp.out[2] = s.mix[2];
        if (v52) { lp.coef[3] = L.att_time_mix_g.data; lp.out[3] = s.mix[3]; }
        cons[0] = &L.att_key; cons[1] = &L.att_value; cons[2] = &L.att_receptance; cons[3] = &L.att_gate;
    }
    (void) layer;
    for (int j = 0; j < lp.n_out; j++) if (cons[j]) { lp.q_out[j] = xq_slot(ctx, T, XQ_ATT_MIX + j, *cons[j]); lp.q_type[j] = cons[j]->type; }
    return lp;
}

bool att_v4(Context * ctx, const Layer & L, const Scratch & s, int T, const float * st_in, float * st_out, const LnMixParams & lp) {
    const int C = ctx->model->n_embed;
    Batch b(T);
    GemvProblem & pr = b.add(L.att_receptance, s.mix[2], s.r, EPI_SIGMOID);
    GemvProblem & pk = b.add(L.att_key, s.mix[0], s.k);
    GemvProblem & pv = b.add(L.att_value, s.mix[1], s.v);
    bool staged = false;
    if (!ln_mix_then(ctx, lp, &staged)) return false;
    if (staged) { xq_bind(pr, lp.q_out[2], lp.q_type[2]); xq_bind(pk, lp.q_out[0], lp.q_type[0]); xq_bind(pv, lp.q_out[1], lp.q_type[1]); }
    if (!run_batch(ctx, b)) return false;
    Wkv4Params wp{};
    wp.k = s.k; wp.v = s.v; wp.r = s.r;
    wp.time_first = L.att_time_first.data; wp.time_decay = L.att_time_decay.data;
    wp.aa_in = st_in + 2 * C; wp.bb_in = st_in + 3 * C; wp.pp_in = st_in + 4 * C;
    wp.aa_out = st_out + 2 * C; wp.bb_out = st_out + 3 * C; wp.pp_out = st_out + 4 * C;
    wp.y = s.y; wp.C = C; wp.T = T;
    wp.q_out = (C % 32 == 0) ? xq_slot(ctx, T, XQ_WKV, L.att_output) : nullptr; wp.q_type = L.att_output.type;
    CUDA_OK(ctx, do_wkv4(ctx, wp));
    return att_output(ctx, L, s, T, wp.q_out);
}

bool att_v5(Context * ctx, const Layer & L, const Scratch & s, int T, const float * st_in, float * st_out, const LnMixParams & lp) {
    const Model & m = *ctx->model;
    const int C = m.n_embed;
    const bool v52 = m.arch_minor >= 2;
    Batch b(T);
    GemvProblem & pr = b.add(L.att_receptance, s.mix[2], s.r);
    GemvProblem & pk = b.add(L.att_key, s.mix[0], s.k);
    GemvProblem & pv = b.add(L.att_value, s.mix[1], s.v);
    GemvProblem * pg = v52 ? &b.add(L.att_gate, s.mix[3], s.g, EPI_SILU) : nullptr;
    bool staged = false;
    if (!ln_mix_then(ctx, lp, &staged)) return false;
    if (staged) {
        xq_bind(pr, lp.q_out[2], lp.q_type[2]); xq_bind(pk, lp.q_out[0], lp.q_type[0]); xq_bind(pv, lp.q_out[1], lp.q_type[1]);
        if (pg) xq_bind(*pg, lp.q_out[3], lp.q_type[3]);
    }
    if (!run_batch(ctx, b)) return false;
    Wkv6Params wp{};
    wp.r = s.r; wp.k = s.k; wp.v = s.v;
    wp.td = L.att_time_decay.data; wp.td_per_token = 0;
    wp.tf = v52 ? L.att_time_faaaa.data : L.att_time_first.data;
    wp.per_head_scalars = v52 ? 0 : 1;
    wp.state_in = st_in + 2 * C; wp.state_out = st_out + 2 * C;
    wp.lnx_w = L.att_ln_x_w.data; wp.lnx_b = L.att_ln_x_b.data;
    wp.g = v52 ? s.g : nullptr;
    wp.y = s.y; wp.eps = 1e-5f; wp.H = m.head_count; wp.S = m.head_size; wp.T = T;
    wp.q_out = (m.head_size % 32 == 0) ? xq_slot(ctx, T, XQ_WKV, L.att_output) : nullptr; wp.q_type = L.att_output.type;
    CUDA_OK(ctx, do_wkv6(ctx, wp));
    return att_output(ctx, L, s, T, wp.q_out);
}

bool att_v6(Context * ctx, const Layer & L, const Scratch & s, int T, const float * st_in, float * st_out, const LnMixParams & lp) {
    const Model & m = *ctx->model;
    const int C = m.n_embed;
    {   // :313-321  tanh(W1 . xxx)
        Batch b(T);
        GemvProblem & p1 = b.add(L.att_maa_w1, s.mix[0], s.lora[0], EPI_TANH);
        bool staged = false;
        if (!ln_mix_then(ctx, lp, &staged)) return false;
        if (staged) xq_bind(p1, lp.q_out[0], lp.q_type[0]);
        if (!run_batch(ctx, b)) return false;
    }
    V6LerpParams vp{};   // :323-346
    vp.w2 = L.att_maa_w2.data; vp.z = s.lora[0]; vp.xx = s.xx; vp.sx = s.sx;
    vp.maa[0] = L.att_maa_w.data; vp.maa[1] = L.att_maa_k.data; vp.maa[2] = L.att_maa_v.data;
    vp.maa[3] = L.att_maa_r.data; vp.maa[4] = L.att_maa_g.data;
    for (int j = 0; j < 5; j++) vp.out[j] = s.mix[1 + j];   // w, k, v, r, g
    vp.C = C; vp.T = T; vp.mix = L.maa_mix;
    {   // staged columns for the five consumers (w, k, v, r, g order of the lerp outputs): all or none
        const DevMatrix * cons[5] = {&L.att_decay_w1, &L.att_key, &L.att_value, &L.att_receptance, &L.att_gate};
        bool all = C % 32 == 0;
        for (int j = 0; j < 5; j++) { vp.q_out[j] = xq_slot(ctx, T, XQ_LERP + j, *cons[j]); vp.q_type[j] = cons[j]->type; all = all && vp.q_out[j]; }
        if (!all) for (int j = 0; j < 5; j++) vp.q_out[j] = nullptr;
    }
    CUDA_OK(ctx, launch_v6_lerp(vp, ctx->stream));
    {   // :349-363
        Batch b(T);
        xq_bind(b.add(L.att_receptance, s.mix[4], s.r), vp.q_out[3], vp.q_type[3]);
        xq_bind(b.add(L.att_key, s.mix[2], s.k), vp.q_out[1], vp.q_type[1]);
        xq_bind(b.add(L.att_value, s.mix[3], s.v), vp.q_out[2], vp.q_type[2]);
        xq_bind(b.add(L.att_gate, s.mix[5], s.g, EPI_SILU), vp.q_out[4], vp.q_type[4]);
        xq_bind(b.add(L.att_decay_w1, s.mix[1], s.lora[1], EPI_TANH), vp.q_out[0], vp.q_type[0]);
        if (!run_batch(ctx, b)) return false;
    }
    {   // :357-367  w = exp(-exp(Wd2 . tanh(..) + time_decay))
        Batch b(T);
        GemvProblem & p = b.add(L.att_decay_w2, s.lora[1], s.w, EPI_BIAS_EXPNEGEXP);
        p.bias = L.att_time_decay.data;
        if (!run_batch(ctx, b)) return false;
    }
    Wkv6Params wp{};   // :370-382
    wp.r = s.r; wp.k = s.k; wp.v = s.v;
    wp.td = s.w; wp.td_per_token = 1; wp.tf = L.att_time_faaaa.data; wp.per_head_scalars = 0;
    wp.state_in = st_in + 2 * C; wp.state_out = st_out + 2 * C;
    wp.lnx_w = L.att_ln_x_w.data; wp.lnx_b = L.att_ln_x_b.data;
    wp.g = s.g; wp.y = s.y; wp.eps = 64e-5f; wp.H = m.head_count; wp.S = m.head_size; wp.T = T;
    wp.q_out = (m.head_size % 32 == 0) ? xq_slot(ctx, T, XQ_WKV, L.att_output) : nullptr; wp.q_type = L.att_output.type;
    CUDA_OK(ctx, do_wkv6(ctx, wp));
    return att_output(ctx, L, s, T, wp.q_out);
}

bool att_v7(Context * ctx, const Layer & L, int layer, const Scratch & s, int T, const float * st_in, float * st_out, const LnMixParams & lp) {
    const Model & m = *ctx->model;
    const int C = m.n_embed;
    const bool first = layer == 0;
    float * v_dst = first ? s.v_first : s.v;
    {   // :415-432, 439, 447 -- first halves of the LoRA pairs. The staged column of a mix output was written for its first consumer;
        // att_v1 shares x_v with att_value only if both multiply the same staged format (in quantised files the LoRA matrices stay fp16)
        Batch b(T);
        GemvProblem * ps[7] = {&b.add(L.att_receptance, s.mix[0], s.r), &b.add(L.att_key, s.mix[2], s.k), &b.add(L.att_value, s.mix[3], v_dst),
                               &b.add(L.att_w1, s.mix[1], s.lora[0], EPI_TANH), &b.add(L.att_a1, s.mix[4], s.lora[1]), &b.add(L.att_g1, s.mix[5], s.lora[2], EPI_SIGMOID),
                               first ? nullptr : &b.add(L.att_v1, s.mix[3], s.lora[3])};
        const int src[7] = {0, 2, 3, 1, 4, 5, 3};      // which mix output each problem reads
        bool staged = false;
        if (!ln_mix_then(ctx, lp, &staged)) return false;
        if (staged) for (int i = 0; i < 7; i++) if (ps[i]) xq_bind(*ps[i], lp.q_out[src[i]], lp.q_type[src[i]]);
        if (!run_batch(ctx, b)) return false;
    }
    {   // second halves
        Batch b(T);
        GemvProblem & pw = b.add(L.att_w2, s.lora[0], s.w, EPI_BIAS_W7);      pw.bias = L.att_w0.data;
        GemvProblem & pa = b.add(L.att_a2, s.lora[1], s.a, EPI_BIAS_SIGMOID); pa.bias = L.att_a0.data;
        b.add(L.att_g2, s.lora[2], s.g);
        if (!first) { GemvProblem & pv = b.add(L.att_v2, s.lora[3], s.vgate, EPI_BIAS_SIGMOID); pv.bias = L.att_v0.data; }
        if (!run_batch(ctx, b)) return false;
    }
    Wkv7Params wp{};   // :433-479
    wp.r = s.r; wp.w = s.w; wp.k = s.k; wp.v = v_dst; wp.a = s.a; wp.g = s.g;
    wp.vgate = first ? nullptr : s.vgate; wp.v_first = s.v_first; wp.v_out = nullptr;
    wp.k_k = L.att_k_k.data; wp.k_a = L.att_k_a.data; wp.r_k = L.att_r_k.data;
    wp.lnx_w = L.att_ln_x_w.data; wp.lnx_b = L.att_ln_x_b.data;
    wp.state_in = st_in + 2 * C; wp.state_out = st_out + 2 * C;
    wp.y = s.y; wp.H = m.head_count; wp.S = m.head_size; wp.T = T;
    wp.q_out = (m.head_size % 32 == 0) ? xq_slot(ctx, T, XQ_WKV, L.att_output) : nullptr; wp.q_type = L.att_output.type;
    CUDA_OK(ctx, ctx->batch_stride ? launch_wkv7_batch(wp, ctx->batch_stride, ctx->stream) : launch_wkv7(wp, ctx->stream));
    return att_output(ctx, L, s, T, wp.q_out);
}

bool ensure_capacity(Context * ctx, int T) {
    if (T <= ctx->capacity_T) return true;
    const Model & m = *ctx->model;
    int cap = ctx->capacity_T ? ctx->capacity_T : 1;
    while (cap < T) cap *= 2;
    if (cap > MAX_TOKENS_PER_PASS) cap = MAX_TOKENS_PER_PASS;
    if (ctx->scratch) { cudaFree(ctx->scratch); ctx->scratch = nullptr; }
    if (ctx->act16) { cudaFree(ctx->act16); ctx->act16 = nullptr; ctx->act16_bytes = 0; }
    if (ctx->tokens) { cudaFree(ctx->tokens); ctx->tokens = nullptr; }
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < 2; i++) {
        if (ctx->tokens_host[i]) { cudaFreeHost(ctx->tokens_host[i]); ctx->tokens_host[i] = nullptr; }
        ctx->slot_used[i] = false;
    }
    for (auto & g : ctx->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); g = Context::GraphSlot(); }
    ctx->capacity_T = 0;
    const size_t n = scratch_floats_for(m, cap);
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&ctx->scratch), n * sizeof(float));
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_CTX | RWKV_ERROR_ALLOC, false, e == cudaSuccess, "Failed to allocate %zu bytes of activation memory: %s", n * sizeof(float), cudaGetErrorString(e));
    e = cudaMalloc(reinterpret_cast<void **>(&ctx->tokens), (size_t) cap * sizeof(int));
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_CTX | RWKV_ERROR_ALLOC, false, e == cudaSuccess, "Failed to allocate the token buffer: %s", cudaGetErrorString(e));
    for (int i = 0; i < 2; i++) {
        e = cudaMallocHost(reinterpret_cast<void **>(&ctx->tokens_host[i]), (size_t) cap * sizeof(int));
        RWKV_CHECK(ctx->sink(), RWKV_ERROR_CTX | RWKV_ERROR_ALLOC, false, e == cudaSuccess, "Failed to allocate pinned token staging: %s", cudaGetErrorString(e));
    }
    if (cap >= BATCH_TC_MIN) {   // fp16 staging for the tensor-core path: up to 8 distinct inputs of max(C, F) x round16(cap)
        Dims d = model_dims(m);
        const size_t kmax = d.F > d.C ? d.F : d.C;
        ctx->act16_bytes = gemm_tc_workspace_bytes(cap, (size_t) GEMV_MAX_PROBLEMS * (kmax * (size_t) ((cap + 15) / 16 * 16) + 128));
        e = cudaMalloc(&ctx->act16, ctx->act16_bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(ctx->act16, 0, GEMM_TC_COUNTER_BYTES, ctx->stream);      // split-K tile counters
        RWKV_CHECK(ctx->sink(), RWKV_ERROR_CTX | RWKV_ERROR_ALLOC, false, e == cudaSuccess, "Failed to allocate fp16 staging: %s", cudaGetErrorString(e));
    }
    ctx->scratch_floats = n;
    ctx->capacity_T = cap;
    return true;
}

// Enqueues one pass (token upload + every kernel) on ctx->stream; pure stream work, so it can be captured.
bool enqueue_pass(Context * ctx, int T, bool want_logits, int phase, int seg) {
    const Model & m = *ctx->model;
    int l0, l1;
    segment_range(ctx, seg, l0, l1);
    g_trace_base = ctx->trace_buf;
    g_trace_next = 0;
    const int C = m.n_embed;
    const Scratch s = carve(m, ctx->scratch, T);
    if (l0 == 0) {
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->tokens, ctx->tokens_host[phase], (size_t) T * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_OK(ctx, launch_embed_ln0(m.emb.data, m.emb.type, m.emb.pitch, ctx->tokens, T, C, m.ln0_w.data, m.ln0_b.data, s.x, ctx->stream));
    }   // a later pipeline stage / layer group finds x (and v_first) already in place
    const size_t per_layer = m.state_floats_per_layer();
    for (int i = l0; i < l1; i++) {
        const Layer & L = m.layers[i];
        const float * st_in = ctx->state_a + (size_t) i * per_layer;
        float * st_out = ctx->state_b + (size_t) i * per_layer;
        const LnMixParams alp = att_ln(ctx, L, i, s, T, st_in, st_out), flp = ffn_ln(ctx, L, s, T, st_in, st_out);
        bool ok;
        switch (m.arch_major) {
            case 7: ok = att_v7(ctx, L, i, s, T, st_in, st_out, alp); break;
            case 6: ok = att_v6(ctx, L, s, T, st_in, st_out, alp); break;
            case 5: ok = att_v5(ctx, L, s, T, st_in, st_out, alp); break;
            default: ok = att_v4(ctx, L, s, T, st_in, st_out, alp); break;
        }
        if (!ok || !ffn(ctx, L, s, T, flp)) return false;
    }
    if (want_logits && l1 == m.n_layer) {   // :705-708 / :851-854  head . LN(x_last; ln_out); a batch context wants every column
        Batch b(ctx->batch_n ? T : 1);
        GemvProblem & p = b.add(m.head, ctx->batch_n ? s.x : s.x + (size_t) (T - 1) * C, ctx->logits);
        p.pro = PRO_LAYERNORM; p.ln_w = m.ln_out_w.data; p.ln_b = m.ln_out_b.data;
        if (!run_batch(ctx, b)) return false;
    }
    if (ctx->trace_buf) {
        ctx->trace_count = g_trace_next;
        for (int i = 0; i < g_trace_next; i++) ctx->trace_names[i] = g_trace_names[i];
    }
    g_trace_base = nullptr;
    return true;
}

// One pass of at most MAX_TOKENS_PER_PASS tokens (state_a -> state_b, then swap) in three steps, so that the overlapped path can
// put copies between the layer groups: begin_pass (capacity, token staging, hand-off in), run_layers per segment, end_pass.
bool begin_pass(Context * ctx, const uint32_t * tokens, int T) {
    if (!ensure_capacity(ctx, T)) return false;
    const int phase = ctx->phase;
    if (ctx->slot_used[phase]) CUDA_OK(ctx, cudaEventSynchronize(ctx->slot_free[phase]));   // pass n-2 has consumed this slot
    const Model & m = *ctx->model;
    if (m.layer_begin == 0) for (int t = 0; t < T; t++) ctx->tokens_host[phase][t] = (int) tokens[t];
    const Scratch hs = carve(m, ctx->scratch, T);
    const size_t ct = (size_t) m.n_embed * (size_t) T * sizeof(float);
    CUDA_OK(ctx, cudaEventRecord(ctx->ev_start, ctx->stream));
    if (ctx->hidden_in) {    // outside the captured graph: the caller's pointers may change from call to call
        CUDA_OK(ctx, cudaMemcpyAsync(hs.x, ctx->hidden_in, ct, cudaMemcpyDeviceToDevice, ctx->stream));
        if (m.arch_major == 7) CUDA_OK(ctx, cudaMemcpyAsync(hs.v_first, ctx->hidden_in + (size_t) m.n_embed * T, ct, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return true;
}

// The kernels of segment `seg` (0 = all resident layers) for the pass begun by begin_pass.
bool run_layers(Context * ctx, int T, bool want_logits, int seg) {
    const Model & m = *ctx->model;
    const int phase = ctx->phase;
    int l0, l1;
    segment_range(ctx, seg, l0, l1);
    want_logits = want_logits && l1 == m.n_layer;        // only the group that ends the model can run the head
    const int slot = Context::slot_index(want_logits, phase, seg);
    Context::GraphSlot * g = (T == 1 && ctx->use_graphs && !ctx->profiling) ? &ctx->graphs[slot] : nullptr;
    if (g && g->exec) {
        CUDA_OK(ctx, cudaGraphLaunch(g->exec, ctx->stream));
        g_kernel_launches += g->launches;
    } else if (g && g->uses >= 1) {
        // second use of this slot: capture the launch sequence once, replay from now on
        const unsigned long long before = g_kernel_launches;
        CUDA_OK(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        const bool ok = enqueue_pass(ctx, T, want_logits, phase, seg);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        RWKV_CHECK(ctx->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, ok && e == cudaSuccess && graph, "CUDA graph capture failed: %s", cudaGetErrorString(e));
        g->launches = g_kernel_launches - before;
        e = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        RWKV_CHECK(ctx->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, e == cudaSuccess, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
        CUDA_OK(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    } else {
        if (g) g->uses++;
        if (!enqueue_pass(ctx, T, want_logits, phase, seg)) return false;
    }
    return true;
}

bool end_pass(Context * ctx, int T, bool want_logits) {
    const Model & m = *ctx->model;
    const int phase = ctx->phase;
    if (ctx->hidden_out) {
        const Scratch hs = carve(m, ctx->scratch, T);
        const size_t ct = (size_t) m.n_embed * (size_t) T * sizeof(float);
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->hidden_out, hs.x, ct, cudaMemcpyDeviceToDevice, ctx->stream));
        if (m.arch_major == 7) CUDA_OK(ctx, cudaMemcpyAsync(ctx->hidden_out + (size_t) m.n_embed * T, hs.v_first, ct, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CUDA_OK(ctx, cudaEventRecord(ctx->ev_stop, ctx->stream));
    CUDA_OK(ctx, cudaEventRecord(ctx->slot_free[phase], ctx->stream));
    ctx->slot_used[phase] = true;
    float * tmp = ctx->state_a; ctx->state_a = ctx->state_b; ctx->state_b = tmp;
    ctx->phase ^= 1;
    ctx->logits_valid = want_logits && m.layer_end == m.n_layer;
    return true;
}

bool forward_pass(Context * ctx, const uint32_t * tokens, int T, bool want_logits) {
    return begin_pass(ctx, tokens, T) && run_layers(ctx, T, want_logits, 0) && end_pass(ctx, T, want_logits);
}

bool ensure_copy_streams(Context * ctx) {
    if (ctx->copy_in) return true;
    CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->pass_begin, cudaEventDisableTiming));
    for (int i = 0; i < Context::MAX_SEGMENTS; i++) {
        CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->seg_in[i], cudaEventDisableTiming));
        CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->seg_out[i], cudaEventDisableTiming));
    }
    return true;
}

}  // namespace

void fill_init_state(const Model & m, float * state) {
    const size_t n = m.state_len();
    memset(state, 0, n * sizeof(float));
    if (m.arch_major >= 5) return;
    const size_t C = m.n_embed;
    for (int l = 0; l < m.n_layer; l++)
        for (size_t c = 0; c < C; c++) state[(size_t) l * 5 * C + 4 * C + c] = -1e30f;
}

Context * create_context(Model * model, ErrorSink sink, int batch_n) {
    Context * ctx = new (std::nothrow) Context();
    RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_ALLOC, nullptr, ctx, "Failed to allocate rwkv_context");
    ctx->model = model;
    model->refcount.fetch_add(1);
    ctx->print_errors = *sink.print;
    { const char * e = getenv("RWKV_B200_OVERLAP"); ctx->overlap_copies = e ? atoi(e) != 0 : OVERLAP_DEFAULT; }
    {   // layer groups of the overlapped host-state path: at least 4 layers each, at most MAX_SEGMENTS groups
        const int n = model->layer_end - model->layer_begin;
        int g = n / 4;
        if (const char * e = getenv("RWKV_B200_SEGMENTS")) g = atoi(e);
        ctx->n_segments = g < 1 ? 1 : (g > Context::MAX_SEGMENTS ? Context::MAX_SEGMENTS : (g > n ? n : g));
    }
    const size_t n = model->state_len();
    const size_t seqs = batch_n > 0 ? (size_t) batch_n : 1;
    if (batch_n > 0) { ctx->batch_n = batch_n; ctx->batch_stride = (long long) n; ctx->overlap_copies = false; }
    bool ok = cudaSetDevice(model->dev.device) == cudaSuccess
        && cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess
        && cudaEventCreate(&ctx->ev_start) == cudaSuccess && cudaEventCreate(&ctx->ev_stop) == cudaSuccess
        && cudaEventCreateWithFlags(&ctx->slot_free[0], cudaEventDisableTiming) == cudaSuccess
        && cudaEventCreateWithFlags(&ctx->slot_free[1], cudaEventDisableTiming) == cudaSuccess
        && cudaMalloc(reinterpret_cast<void **>(&ctx->state_a), seqs * n * sizeof(float)) == cudaSuccess
        && cudaMalloc(reinterpret_cast<void **>(&ctx->state_b), seqs * n * sizeof(float)) == cudaSuccess
        && cudaMalloc(reinterpret_cast<void **>(&ctx->state_init), n * sizeof(float)) == cudaSuccess
        && cudaMalloc(reinterpret_cast<void **>(&ctx->logits), seqs * (size_t) model->n_vocab * sizeof(float)) == cudaSuccess;
    if (ok && batch_n == 0) {      // staged-column hand-off slots (every handed-off vector has n_embed elements, at most 2 bytes each) 
        ctx->xq_slot_bytes = ((size_t) 2 * model->n_embed + 64 + 255) / 256 * 256;
        ok = cudaMalloc(reinterpret_cast<void **>(&ctx->xq), (size_t) XQ_SLOTS * ctx->xq_slot_bytes) == cudaSuccess;
    }
    if (ok) {
        std::vector<float> init(n);
        fill_init_state(*model, init.data());
        ok = cudaMemcpy(ctx->state_init, init.data(), n * sizeof(float), cudaMemcpyHostToDevice) == cudaSuccess;
        for (size_t q = 0; ok && batch_n > 0 && q < seqs; q++)
            ok = cudaMemcpy(ctx->state_a + q * n, ctx->state_init, n * sizeof(float), cudaMemcpyDeviceToDevice) == cudaSuccess;
    }
    if (!ok) {
        cudaError_t e = cudaGetLastError();
        destroy_context(ctx);
        RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_ALLOC, nullptr, false, "Failed to set up the context on the device: %s", cudaGetErrorString(e));
    }
    return ctx;
}

void destroy_context(Context * ctx) {
    if (!ctx) return;
    for (Context * s : ctx->stages) destroy_context(s);
    ctx->stages.clear();
    if (PipeGroup * g = ctx->group) {
        ctx->group = nullptr;
        if (g->refs.fetch_sub(1) == 1) {
            for (size_t r = 0; r < g->streams.size(); r++) { cudaSetDevice(g->devices[r]); cudaStreamSynchronize(g->streams[r]); cudaStreamDestroy(g->streams[r]); }
            delete g;
        }
    }
    Model * model = ctx->model;
    if (model) cudaSetDevice(model->dev.device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->state_a); cudaFree(ctx->state_b); cudaFree(ctx->state_init); cudaFree(ctx->logits);
    cudaFree(ctx->tokens); cudaFree(ctx->scratch); cudaFree(ctx->trace_buf); cudaFree(ctx->act16); cudaFree(ctx->xq);
    for (int i = 0; i < 2; i++) {
        if (ctx->tokens_host[i]) cudaFreeHost(ctx->tokens_host[i]);
        if (ctx->slot_free[i]) cudaEventDestroy(ctx->slot_free[i]);
    }
    for (auto & g : ctx->graphs) if (g.exec) cudaGraphExecDestroy(g.exec);
    if (ctx->copy_pool) destroy_copy_pool(ctx->copy_pool);
    if (ctx->bounce_in) { cudaFreeHost(ctx->bounce_in); cudaFreeHost(ctx->bounce_out); cudaFreeHost(ctx->bounce_logits); }
    for (int i = 0; i < Context::MAX_SEGMENTS; i++) if (ctx->seg_d2h[i]) cudaEventDestroy(ctx->seg_d2h[i]);
    if (ctx->copy_in) { cudaStreamSynchronize(ctx->copy_in); cudaStreamDestroy(ctx->copy_in); }
    if (ctx->copy_out) { cudaStreamSynchronize(ctx->copy_out); cudaStreamDestroy(ctx->copy_out); }
    if (ctx->pass_begin) cudaEventDestroy(ctx->pass_begin);
    if (ctx->pipe_done) cudaEventDestroy(ctx->pipe_done);
    for (int i = 0; i < Context::MAX_SEGMENTS; i++) { if (ctx->seg_in[i]) cudaEventDestroy(ctx->seg_in[i]); if (ctx->seg_out[i]) cudaEventDestroy(ctx->seg_out[i]); }
    cudaFree(ctx->sample_token); cudaFree(ctx->sample_scratch); cudaFree(ctx->bias_ids); cudaFree(ctx->bias_values);
    if (ctx->sample_token_host) cudaFreeHost(ctx->sample_token_host);
    for (auto & r : ctx->prof) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    if (ctx->ev_stop) cudaEventDestroy(ctx->ev_stop);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (model && model->refcount.fetch_sub(1) == 1) delete model;
    delete ctx;
}

bool upload_state(Context * ctx, const float * state_in) {
    const size_t bytes = ctx->model->state_len() * sizeof(float);
    // cudaMemcpyDefault: state_in / state_out / logits_out may be host memory (the rwkv.h contract) or device memory of this process
    // (a CUDA tensor handed over by the Python wrapper): unified addressing tells them apart
    if (state_in) CUDA_OK(ctx, cudaMemcpyAsync(ctx->state_a, state_in, bytes, cudaMemcpyDefault, ctx->stream));
    else CUDA_OK(ctx, cudaMemcpyAsync(ctx->state_a, ctx->state_init, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return true;
}

bool download_outputs(Context * ctx, float * state_out, float * logits_out) {
    if (state_out) CUDA_OK(ctx, cudaMemcpyAsync(state_out, ctx->state_a, ctx->model->state_len() * sizeof(float), cudaMemcpyDefault, ctx->stream));
    if (logits_out) CUDA_OK(ctx, cudaMemcpyAsync(logits_out, ctx->logits, (size_t) ctx->model->n_vocab * sizeof(float), cudaMemcpyDefault, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return true;
}

// ---- pageable caller buffers ----------------------------------------------------------------------------------------------------
// The reference's bindings hand over ordinary (pageable) memory (python/rwkv_cpp/rwkv_cpp_model.py:330-351: fresh torch / numpy buffers per
// call). A cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread, synchronously and at ~6 GB/s: the
// eight slice copies of a 7B state took 5.8 ms before the first kernel was even enqueued (121 tok/s against 303 with pinned buffers,
// profiles/r2_c8_*.json). Registering the caller's pages (cudaHostRegister) would be as fast as pinned memory, but a registration
// outlives a free(): the next buffer malloc() places at the same address would be DMA'd from the OLD pages. So the state travels
// through pinned bounce buffers owned by the context, copied by two helper threads (one per direction) slice by slice, in step with
// the layer groups: memcpy of slice g+1 || H2D of slice g || kernels of group g-1 ... and the mirror image on the way out.
static bool is_pageable_host(const void * p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
static bool bounce_enabled() {
    static const bool on = getenv("RWKV_B200_NO_BOUNCE") == nullptr;
    return on;
}
// states below this size go the plain way (two helper threads cost ~0.1 ms); rwkv_b200_set_bounce_min_bytes moves the bar (tests)
std::atomic<size_t> g_bounce_min_bytes{(size_t) 4 << 20};
CopyPool * new_copy_pool();
static bool ensure_bounce(Context * ctx) {
    if (ctx->bounce_in) return true;
    const size_t bytes = ctx->model->state_len() * sizeof(float);
    if (cudaMallocHost(reinterpret_cast<void **>(&ctx->bounce_in), bytes) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void **>(&ctx->bounce_out), bytes) != cudaSuccess ||
        cudaMallocHost(reinterpret_cast<void **>(&ctx->bounce_logits), (size_t) ctx->model->n_vocab * sizeof(float)) != cudaSuccess) {
        cudaGetLastError();
        if (ctx->bounce_in) cudaFreeHost(ctx->bounce_in);
        if (ctx->bounce_out) cudaFreeHost(ctx->bounce_out);
        ctx->bounce_in = ctx->bounce_out = ctx->bounce_logits = nullptr;
        return false;                      // not an error: the caller falls back to plain copies
    }
    for (int i = 0; i < Context::MAX_SEGMENTS; i++)
        if (cudaEventCreateWithFlags(&ctx->seg_d2h[i], cudaEventDisableTiming) != cudaSuccess) return false;
    ctx->copy_pool = new_copy_pool();
    return ctx->copy_pool != nullptr;
}
}  // namespace rwkv (the pool type is named in engine.h)
namespace rwkv {

// The helper threads of the bounce path: COPY_LANES per direction, created with the bounce buffers and parked on a condition variable
// between passes (spawning eight threads per token would cost more than the copies). One job = one pass: lane k of a direction copies
// the k-th part of every slice, so slices complete in order at the combined memcpy rate of the lanes (one thread moves ~6 GB/s into
// pinned memory on the bench host: 5.8 ms for the 34.6 MB state of a 7B model; profiles/r2_c13_bench_7b.json, 157 tok/s).
struct CopyPool {
    static constexpr int LANES = 4, N = 2 * LANES;
    std::thread th[N];
    std::mutex m;
    std::condition_variable cv;
    unsigned long long generation = 0;
    bool quit = false;
    std::function<void(int)> job;
    std::atomic<int> running{0};
    CopyPool() {
        for (int k = 0; k < N; k++) th[k] = std::thread([this, k] {
            unsigned long long seen = 0;
            for (;;) {
                std::function<void(int)> f;
                {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return quit || generation != seen; });
                    if (quit) return;
                    seen = generation;
                    f = job;
                }
                f(k);
                running.fetch_sub(1, std::memory_order_release);
            }
        });
    }
    void start(std::function<void(int)> f) {
        { std::lock_guard<std::mutex> lk(m); job = std::move(f); running.store(N); generation++; }
        cv.notify_all();
    }
    void wait() { while (running.load(std::memory_order_acquire) > 0) std::this_thread::yield(); }
    ~CopyPool() {
        wait();
        { std::lock_guard<std::mutex> lk(m); quit = true; }
        cv.notify_all();
        for (auto & t : th) if (t.joinable()) t.join();
    }
};
void destroy_copy_pool(CopyPool * p) { delete p; }
CopyPool * new_copy_pool() { return new (std::nothrow) CopyPool(); }

namespace {
// Progress of one pass's copies; waits for the lanes whatever way the pass ends.
struct BouncePass {
    CopyPool * pool = nullptr;
    std::atomic<int> in_parts[Context::MAX_SEGMENTS];     // lanes that have copied their part of slice g into bounce_in
    std::atomic<int> out_posted{0};                        // slices whose D2H copy + event have been enqueued
    std::atomic<int> out_arrived{0};                       // slices that have landed in bounce_out
    std::atomic<bool> abort{false};
    BouncePass() { for (auto & a : in_parts) a.store(0); }
    ~BouncePass() { if (pool) { abort.store(true); pool->wait(); } }
};
}  // namespace

// One pass with the caller's host state pipelined against the layer groups: H2D of group g+1 and D2H of group g-1 run on their
// own streams while group g computes (the state layout is layer-major, rwkv_graph.inc:545-606, so a group is one contiguous
// slice). state_in == NULL starts from the init image (a device copy); state_out / logits_out may be NULL.
static bool eval_host_overlapped(Context * ctx, const uint32_t * tokens, int T, const float * state_in, bool resident_in, float * state_out, float * logits_out) {
    const Model & m = *ctx->model;
    const int G = ctx->n_segments;
    if (!ensure_copy_streams(ctx)) return false;
    const size_t per_layer = m.state_floats_per_layer();
    const size_t state_bytes = m.state_len() * sizeof(float);
    // pageable caller memory takes the bounce path (worth waking eight threads from a few MB on)
    const bool big = state_bytes >= g_bounce_min_bytes.load() && bounce_enabled();
    bool b_in = big && state_in && is_pageable_host(state_in);
    bool b_out = big && state_out && is_pageable_host(state_out);
    if ((b_in || b_out) && !ensure_bounce(ctx)) b_in = b_out = false;
    const bool b_log = b_out && logits_out && is_pageable_host(logits_out);
    const int dev = m.dev.device;
    auto slice = [&](int g, size_t & off, size_t & cnt) { int l0, l1; segment_range(ctx, g, l0, l1); off = (size_t) l0 * per_layer; cnt = (size_t) (l1 - l0) * per_layer; };
    constexpr int LANES = CopyPool::LANES;
    auto part = [&](int g, int lane, size_t & off, size_t & cnt) {      // lane's share of slice g, in whole 64-byte lines
        size_t o, c; slice(g, o, c);
        const size_t per = ((c + LANES - 1) / LANES + 15) & ~(size_t) 15;
        const size_t b = std::min(c, per * (size_t) lane), e = std::min(c, per * (size_t) (lane + 1));
        off = o + b; cnt = e - b;
    };
    BouncePass bp;
    if (b_in || b_out) {
        bp.pool = ctx->copy_pool;
        bp.pool->start([&, G, dev, b_in, b_out](int k) {
            const bool inward = k < LANES;
            const int lane = inward ? k : k - LANES;
            if (inward) {
                if (!b_in) return;
                for (int g = 1; g <= G && !bp.abort.load(); g++) {
                    size_t off, cnt; part(g, lane, off, cnt);
                    if (cnt) memcpy(ctx->bounce_in + off, state_in + off, cnt * sizeof(float));
                    bp.in_parts[g - 1].fetch_add(1, std::memory_order_release);
                }
            } else {
                if (!b_out) return;
                if (lane == 0) cudaSetDevice(dev);
                for (int g = 1; g <= G; g++) {
                    if (lane == 0) {      // one lane waits for the DMA, the others for that lane
                        while (bp.out_posted.load(std::memory_order_acquire) < g) { if (bp.abort.load()) return; std::this_thread::yield(); }
                        if (cudaEventSynchronize(ctx->seg_d2h[g - 1]) != cudaSuccess) { bp.abort.store(true); return; }
                        bp.out_arrived.store(g, std::memory_order_release);
                    } else {
                        while (bp.out_arrived.load(std::memory_order_acquire) < g) { if (bp.abort.load()) return; std::this_thread::yield(); }
                    }
                    size_t off, cnt; part(g, lane, off, cnt);
                    if (cnt) memcpy(state_out + off, ctx->bounce_out + off, cnt * sizeof(float));
                }
            }
        });
    }
    // nothing of this pass may start before everything enqueued earlier on the context's stream has finished with the state
    CUDA_OK(ctx, cudaEventRecord(ctx->pass_begin, ctx->stream));
    CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_in, ctx->pass_begin, 0));
    if (state_in && !b_in) {
        for (int g = 1; g <= G; g++) {
            size_t off, cnt; slice(g, off, cnt);
            CUDA_OK(ctx, cudaMemcpyAsync(ctx->state_a + off, state_in + off, cnt * sizeof(float), cudaMemcpyDefault, ctx->copy_in));
            CUDA_OK(ctx, cudaEventRecord(ctx->seg_in[g - 1], ctx->copy_in));
        }
    } else if (!state_in && !resident_in) {
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->state_a, ctx->state_init, m.state_len() * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    }
    if (!begin_pass(ctx, tokens, T)) return false;
    const bool want_logits = logits_out != nullptr;
    for (int g = 1; g <= G; g++) {
        if (b_in) {      // the slice is in pinned memory once every lane has delivered its part
            while (bp.in_parts[g - 1].load(std::memory_order_acquire) < LANES) std::this_thread::yield();
            size_t off, cnt; slice(g, off, cnt);
            CUDA_OK(ctx, cudaMemcpyAsync(ctx->state_a + off, ctx->bounce_in + off, cnt * sizeof(float), cudaMemcpyHostToDevice, ctx->copy_in));
            CUDA_OK(ctx, cudaEventRecord(ctx->seg_in[g - 1], ctx->copy_in));
        }
        if (state_in) CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->seg_in[g - 1], 0));
        if (!run_layers(ctx, T, want_logits, g)) return false;
        CUDA_OK(ctx, cudaEventRecord(ctx->seg_out[g - 1], ctx->stream));
        if (b_out && g > 1) {      // one group behind: the copy-out stream never waits for something that has not been enqueued
            size_t off, cnt; slice(g - 1, off, cnt);
            // the new state of group g-1 is in state_b until end_pass swaps the buffers
            CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->seg_out[g - 2], 0));
            CUDA_OK(ctx, cudaMemcpyAsync(ctx->bounce_out + off, ctx->state_b + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out));
            CUDA_OK(ctx, cudaEventRecord(ctx->seg_d2h[g - 2], ctx->copy_out));
            bp.out_posted.store(g - 1, std::memory_order_release);
        }
    }
    if (!end_pass(ctx, T, want_logits)) return false;      // the new state is ctx->state_a from here on
    if (logits_out) CUDA_OK(ctx, cudaMemcpyAsync(b_log ? ctx->bounce_logits : logits_out, ctx->logits, (size_t) m.n_vocab * sizeof(float), cudaMemcpyDefault, ctx->stream));
    if (state_out && b_out) {
        size_t off, cnt; slice(G, off, cnt);
        CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->seg_out[G - 1], 0));
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->bounce_out + off, ctx->state_a + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, ctx->copy_out));
        CUDA_OK(ctx, cudaEventRecord(ctx->seg_d2h[G - 1], ctx->copy_out));
        bp.out_posted.store(G, std::memory_order_release);
    } else if (state_out) {
        for (int g = 1; g <= G; g++) {
            size_t off, cnt; slice(g, off, cnt);
            CUDA_OK(ctx, cudaStreamWaitEvent(ctx->copy_out, ctx->seg_out[g - 1], 0));
            CUDA_OK(ctx, cudaMemcpyAsync(state_out + off, ctx->state_a + off, cnt * sizeof(float), cudaMemcpyDefault, ctx->copy_out));
        }
    }
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    if (b_log) memcpy(logits_out, ctx->bounce_logits, (size_t) m.n_vocab * sizeof(float));
    if (state_out) CUDA_OK(ctx, cudaStreamSynchronize(ctx->copy_out));
    if (bp.pool) { bp.pool->wait(); bp.pool = nullptr; }      // the last slices have been copied to the caller
    return !bp.abort.load();
}

static bool can_overlap(const Context * ctx) {
    const Model & m = *ctx->model;
    return ctx->overlap_copies && ctx->n_segments > 1 && m.layer_begin == 0 && m.layer_end == m.n_layer && !ctx->profiling;
}

bool eval_host(Context * ctx, const uint32_t * tokens, size_t T, const float * state_in, float * state_out, float * logits_out) {
    CUDA_OK(ctx, cudaSetDevice(ctx->model->dev.device));
    if (can_overlap(ctx) && T <= (size_t) MAX_TOKENS_PER_PASS && (state_in || state_out))
        return eval_host_overlapped(ctx, tokens, (int) T, state_in, false, state_out, logits_out);
    return upload_state(ctx, state_in) && forward(ctx, tokens, T, logits_out != nullptr) && download_outputs(ctx, state_out, logits_out);
}

bool eval_host_chunks(Context * ctx, const uint32_t * tokens, size_t T, size_t chunk, const float * state_in, float * state_out, float * logits_out) {
    CUDA_OK(ctx, cudaSetDevice(ctx->model->dev.device));
    const bool overlap = can_overlap(ctx) && chunk <= (size_t) MAX_TOKENS_PER_PASS && (state_in || state_out);
    if (!overlap && !upload_state(ctx, state_in)) return false;
    // Same chunk boundaries as the reference loop (rwkv_eval.inc:179-218); the state never leaves HBM between chunks, logits are
    // computed for the final chunk only. With overlap the first chunk takes the caller's state group by group while it computes and
    // the last one hands the new state back the same way.
    size_t off = 0;
    while (off < T) {
        const size_t n = T - off < chunk ? T - off : chunk;
        const bool first = off == 0, last = off + n == T;
        if (overlap && (first || last)) {
            if (!eval_host_overlapped(ctx, tokens + off, (int) n, first ? state_in : nullptr, !first, last ? state_out : nullptr, last ? logits_out : nullptr)) return false;
        } else if (!forward(ctx, tokens + off, n, last && logits_out != nullptr)) {
            return false;
        }
        off += n;
    }
    return overlap ? true : download_outputs(ctx, state_out, logits_out);
}

bool forward(Context * ctx, const uint32_t * tokens, size_t T, bool want_logits) {
    CUDA_OK(ctx, cudaSetDevice(ctx->model->dev.device));
    size_t done = 0;
    while (done < T) {
        const size_t n = (T - done < (size_t) MAX_TOKENS_PER_PASS) ? T - done : (size_t) MAX_TOKENS_PER_PASS;
        const bool last = done + n == T;
        if (!forward_pass(ctx, tokens + done, (int) n, want_logits && last)) return false;
        done += n;
    }
    return true;
}

bool batch_set_state(Context * ctx, int seq, const float * state_in) {
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, ctx->batch_n > 0 && seq >= 0 && seq < ctx->batch_n, "Not a batch context or sequence index out of range");
    CUDA_OK(ctx, cudaSetDevice(ctx->model->dev.device));
    const size_t n = ctx->model->state_len();
    float * dst = ctx->state_a + (size_t) seq * n;
    if (state_in) CUDA_OK(ctx, cudaMemcpyAsync(dst, state_in, n * sizeof(float), cudaMemcpyDefault, ctx->stream));
    else CUDA_OK(ctx, cudaMemcpyAsync(dst, ctx->state_init, n * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
    return true;
}

bool batch_get_state(Context * ctx, int seq, float * state_out) {
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, ctx->batch_n > 0 && seq >= 0 && seq < ctx->batch_n && state_out, "Not a batch context, sequence index out of range or NULL buffer");
    CUDA_OK(ctx, cudaSetDevice(ctx->model->dev.device));
    const size_t n = ctx->model->state_len();
    CUDA_OK(ctx, cudaMemcpyAsync(state_out, ctx->state_a + (size_t) seq * n, n * sizeof(float), cudaMemcpyDefault, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return true;
}

bool batch_eval(Context * ctx, const uint32_t * tokens, bool want_logits) {
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, ctx->batch_n > 0 && tokens, "Not a batch context or NULL tokens");
    CUDA_OK(ctx, cudaSetDevice(ctx->model->dev.device));
    return forward_pass(ctx, tokens, ctx->batch_n, want_logits);
}

bool batch_get_logits(Context * ctx, int seq, float * logits_out) {
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, ctx->batch_n > 0 && seq >= 0 && seq < ctx->batch_n && logits_out, "Not a batch context, sequence index out of range or NULL buffer");
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, ctx->logits_valid, "The last batch evaluation skipped the head");
    CUDA_OK(ctx, cudaSetDevice(ctx->model->dev.device));
    const size_t V = (size_t) ctx->model->n_vocab;
    CUDA_OK(ctx, cudaMemcpyAsync(logits_out, ctx->logits + (size_t) seq * V, V * sizeof(float), cudaMemcpyDefault, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    return true;
}

bool sample_token(Context * ctx, float temperature, float top_p, double u, const uint32_t * bias_ids, const float * bias_values, size_t n_bias, uint32_t * token_out) {
    const Model & m = *ctx->model;
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, token_out != nullptr, "token_out is NULL");
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, ctx->batch_n == 0, "Sampling on a batch context is not supported");
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, temperature >= 0.0f, "temperature must be >= 0");            // sampling.py:20-21
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, top_p >= 0.0f && top_p <= 1.0f, "top_p must be in [0, 1]");    // sampling.py:22-23
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, u >= 0.0 && u < 1.0, "u must be in [0, 1)");
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, n_bias == 0 || (bias_ids && bias_values), "logit bias arrays are NULL");
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, ctx->logits_valid, "No logits to sample from: the last evaluation skipped the head");
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, m.n_vocab <= 65536, "On-device sampling supports n_vocab <= 65536");
    CUDA_OK(ctx, cudaSetDevice(m.dev.device));
    if (!ctx->sample_token) {
        CUDA_OK(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->sample_token), 2 * sizeof(uint32_t)));
        CUDA_OK(ctx, cudaMallocHost(reinterpret_cast<void **>(&ctx->sample_token_host), 2 * sizeof(uint32_t)));
    }
    for (size_t i = 0; i < n_bias; i++)
        RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, bias_ids[i] < (uint32_t) m.n_vocab, "logit bias id %u is out of range", bias_ids[i]);
    if (n_bias > 0) {
        if (!ctx->sample_scratch) CUDA_OK(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->sample_scratch), (size_t) m.n_vocab * sizeof(float)));
        if (n_bias > ctx->bias_capacity) {
            CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
            cudaFree(ctx->bias_ids); cudaFree(ctx->bias_values);
            ctx->bias_ids = nullptr; ctx->bias_values = nullptr; ctx->bias_capacity = 0;
            CUDA_OK(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->bias_ids), n_bias * sizeof(uint32_t)));
            CUDA_OK(ctx, cudaMalloc(reinterpret_cast<void **>(&ctx->bias_values), n_bias * sizeof(float)));
            ctx->bias_capacity = n_bias;
        }
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->bias_ids, bias_ids, n_bias * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_OK(ctx, cudaMemcpyAsync(ctx->bias_values, bias_values, n_bias * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    }
    SampleParams sp{};
    sp.logits = ctx->logits; sp.n_vocab = m.n_vocab; sp.temperature = temperature; sp.top_p = top_p; sp.u = u;
    sp.bias_ids = ctx->bias_ids; sp.bias_values = ctx->bias_values; sp.n_bias = (int) n_bias;
    sp.scratch = ctx->sample_scratch; sp.token_out = ctx->sample_token; sp.prob_out = nullptr;
    CUDA_OK(ctx, launch_sample(sp, ctx->stream));
    CUDA_OK(ctx, cudaMemcpyAsync(ctx->sample_token_host, ctx->sample_token, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
    *token_out = ctx->sample_token_host[0];
    return true;
}

bool pipe_ensure_box(Context * ctx) {
    Model & m = *ctx->model;
    if (m.link.box) return true;
    CUDA_OK(ctx, cudaSetDevice(m.dev.device));
    m.link.slot_floats = stage_hidden_len(m, (size_t) MAX_TOKENS_PER_PASS);
    m.link.bytes = sizeof(PipeBox) + (size_t) PIPE_SLOTS * m.link.slot_floats * sizeof(float);
    CUDA_OK(ctx, cudaMalloc(reinterpret_cast<void **>(&m.link.box), m.link.bytes));
    CUDA_OK(ctx, cudaMemset(m.link.box, 0, sizeof(PipeBox)));
    CUDA_OK(ctx, cudaMalloc(reinterpret_cast<void **>(&m.link.counters), 8 * sizeof(unsigned long long)));
    CUDA_OK(ctx, cudaMemset(m.link.counters, 0, 8 * sizeof(unsigned long long)));
    CUDA_OK(ctx, cudaDeviceSynchronize());
    return true;
}

bool pipe_forward(Context * ctx, const uint32_t * tokens, size_t T, bool want_logits, cudaStream_t stream) {
    Model & m = *ctx->model;
    const bool first = m.layer_begin == 0, last = m.layer_end == m.n_layer;
    RWKV_CHECK(ctx->sink(), RWKV_ERROR_ARGS, false, m.link.box && (first || m.link.prev) && (last || m.link.next), "The stage is not connected to its neighbours");
    CUDA_OK(ctx, cudaSetDevice(m.dev.device));
    cudaStream_t own = ctx->stream;
    if (stream) ctx->stream = stream;
    bool ok = begin_pass(ctx, tokens, (int) T);
    if (ok) {
        const Scratch hs = carve(m, ctx->scratch, (int) T);
        const size_t n = (size_t) m.n_embed * T, nv = m.arch_major == 7 ? n : 0;
        cudaError_t e = cudaSuccess;
        if (!first) e = launch_pipe_recv(m.link.box, m.link.prev, m.link.counters, m.link.slot_floats, hs.x, n, hs.v_first, nv, ctx->stream);
        ok = e == cudaSuccess && run_layers(ctx, (int) T, want_logits && last, 0);
        if (ok && !last) e = launch_pipe_send(m.link.box, m.link.next, m.link.counters, m.link.slot_floats, hs.x, n, hs.v_first, nv, ctx->stream);
        ok = ok && e == cudaSuccess && end_pass(ctx, (int) T, want_logits && last);
        if (e != cudaSuccess) { ctx->last_error |= RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED; if (ctx->print_errors) fprintf(stderr, "pipeline hand-off launch failed: %s\n", cudaGetErrorString(e)); }
    }
    ctx->stream = own;
    return ok;
}

// ---- in-process pipeline behind rwkv.h ---------------------------------------------------------------------------
namespace {

// contiguous min-max partition of n_layer layers over `world` stages, the last one carrying `head_layers` extra (pipeline.py:
// stage_layers_balanced is the same recurrence)
std::vector<int> balanced_cuts(int n_layer, int world, double head_layers) {
    const double INF = 1e300;
    std::vector<std::vector<double>> best((size_t) world, std::vector<double>((size_t) n_layer + 1, INF));
    std::vector<std::vector<int>> cut((size_t) world, std::vector<int>((size_t) n_layer + 1, 0));
    auto cost = [&](int st, int a, int b) { return (double) (b - a) + (st == world - 1 ? head_layers : 0.0); };
    for (int e = 1; e <= n_layer; e++) best[0][(size_t) e] = cost(0, 0, e);
    for (int st = 1; st < world; st++)
        for (int e = st + 1; e <= n_layer; e++)
            for (int a = st; a < e; a++) {
                const double v = std::max(best[(size_t) st - 1][(size_t) a], cost(st, a, e));
                if (v < best[(size_t) st][(size_t) e] - 1e-12) { best[(size_t) st][(size_t) e] = v; cut[(size_t) st][(size_t) e] = a; }
            }
    std::vector<int> bounds((size_t) world + 1, 0);
    bounds[(size_t) world] = n_layer;
    for (int st = world - 1, e = n_layer; st >= 1; st--) { e = cut[(size_t) st][(size_t) e]; bounds[(size_t) st] = e; }
    return bounds;
}

Context * stage_of(Context * head, size_t r) { return r == 0 ? head : head->stages[r - 1]; }

bool connect_stages(Context * head, ErrorSink sink) {
    const size_t n = head->stages.size() + 1;
    for (size_t r = 0; r < n; r++) if (!pipe_ensure_box(stage_of(head, r))) return false;
    for (size_t r = 0; r < n; r++) {
        Model & m = *stage_of(head, r)->model;
        for (int dir = 0; dir < 2; dir++) {
            if ((dir == 0 && r == 0) || (dir == 1 && r + 1 == n)) continue;
            Model & o = *stage_of(head, dir == 0 ? r - 1 : r + 1)->model;
            if (o.dev.device != m.dev.device) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, m.dev.device, o.dev.device);
                RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, can, "Device %d cannot access device %d's memory", m.dev.device, o.dev.device);
                cudaSetDevice(m.dev.device);
                const cudaError_t e = cudaDeviceEnablePeerAccess(o.dev.device, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, e == cudaSuccess, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
            }
            (dir == 0 ? m.link.prev : m.link.next) = o.link.box;
        }
    }
    return true;
}

}  // namespace

Context * create_pipeline(const char * path, const std::vector<int> & devices, ErrorSink sink) {
    ModelFile mf;
    if (!scan_model_file(path, mf, sink)) return nullptr;
    const int n_layer = (int) mf.header.n_layer, world = (int) devices.size();
    RWKV_CHECK(sink, RWKV_ERROR_ARGS, nullptr, world >= 2 && world <= n_layer, "A pipeline needs 2 .. n_layer (%d) stages, got %d", n_layer, world);
    double layer_bytes = 0, head_bytes = 0;
    for (const TensorInfo & t : mf.tensors) {
        if (t.name == "head.weight") head_bytes += (double) t.nbytes;
        else if (t.name.compare(0, 7, "blocks.") == 0) layer_bytes += (double) t.nbytes;
    }
    layer_bytes /= n_layer;
    const std::vector<int> bounds = balanced_cuts(n_layer, world, layer_bytes > 0 ? head_bytes / layer_bytes : 0.0);
    Context * head = nullptr;
    PipeGroup * group = new (std::nothrow) PipeGroup();
    RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_ALLOC, nullptr, group, "Failed to allocate the pipeline group");
    group->devices = devices;
    bool ok = true;
    for (int r = 0; r < world && ok; r++) {
        Model * m = load_model(path, devices[(size_t) r], bounds[(size_t) r], bounds[(size_t) r + 1], sink);
        Context * c = m ? create_context(m, sink) : nullptr;
        if (!c) { ok = false; break; }
        c->overlap_copies = false;
        if (r == 0) { head = c; head->group = group; group->refs.fetch_add(1); } else head->stages.push_back(c);
        cudaStream_t s = nullptr;
        ok = cudaSetDevice(devices[(size_t) r]) == cudaSuccess && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
        group->streams.push_back(s);
    }
    ok = ok && connect_stages(head, sink);
    if (!ok) {
        if (head) destroy_context(head); else delete group;
        return nullptr;
    }
    return head;
}

Context * clone_pipeline(Context * src, ErrorSink sink) {
    Context * head = create_context(src->model, sink);
    if (!head) return nullptr;
    head->overlap_copies = false;
    head->group = src->group;
    src->group->refs.fetch_add(1);
    for (Context * s : src->stages) {
        Context * c = create_context(s->model, sink);
        if (!c) { destroy_context(head); return nullptr; }
        c->overlap_copies = false;
        head->stages.push_back(c);
    }
    return head;
}

bool pipeline_eval_host(Context * head, const uint32_t * tokens, size_t T, size_t chunk, const float * state_in, float * state_out, float * logits_out) {
    PipeGroup & g = *head->group;
    const size_t n = head->stages.size() + 1;
    const size_t per_layer = head->model->state_floats_per_layer();
    {
        std::lock_guard<std::mutex> lock(g.order);      // the passes of one evaluation enter every link back to back
        for (size_t r = 0; r < n; r++) {                // every stage takes ITS slice of the caller's state (layer-major ABI layout)
            Context * c = stage_of(head, r);
            const Model & m = *c->model;
            CUDA_OK(head, cudaSetDevice(m.dev.device));
            const size_t off = (size_t) m.layer_begin * per_layer, cnt = (size_t) (m.layer_end - m.layer_begin) * per_layer;
            CUDA_OK(head, cudaMemcpyAsync(c->state_a + off, state_in ? state_in + off : c->state_init + off, cnt * sizeof(float), cudaMemcpyDefault, g.streams[r]));
        }
        // chunk boundaries as the reference loop (rwkv_eval.inc:179-218), each cut further into passes the kernels take
        size_t off = 0;
        while (off < T) {
            size_t nchunk = T - off;
            if (chunk && nchunk > chunk) nchunk = chunk;
            size_t done = 0;
            while (done < nchunk) {
                const size_t np = nchunk - done < (size_t) MAX_TOKENS_PER_PASS ? nchunk - done : (size_t) MAX_TOKENS_PER_PASS;
                const bool final_pass = off + done + np == T;
                for (size_t r = 0; r < n; r++) {
                    Context * c = stage_of(head, r);
                    if (!pipe_forward(c, tokens + off + done, np, final_pass && logits_out != nullptr, g.streams[r])) { head->last_error |= c->last_error; return false; }
                }
                done += np;
            }
            off += nchunk;
        }
        for (size_t r = 0; r < n; r++) {
            Context * c = stage_of(head, r);
            const Model & m = *c->model;
            CUDA_OK(head, cudaSetDevice(m.dev.device));
            const size_t so = (size_t) m.layer_begin * per_layer, cnt = (size_t) (m.layer_end - m.layer_begin) * per_layer;
            if (state_out) CUDA_OK(head, cudaMemcpyAsync(state_out + so, c->state_a + so, cnt * sizeof(float), cudaMemcpyDefault, g.streams[r]));
            if (logits_out && r + 1 == n) CUDA_OK(head, cudaMemcpyAsync(logits_out, c->logits, (size_t) m.n_vocab * sizeof(float), cudaMemcpyDefault, g.streams[r]));
            // completion marker of THIS evaluation on stage r. Waiting on the stream itself (cudaStreamSynchronize) outside the lock
            // would be illegal whenever another host thread is capturing its stage graph on the same stream at that moment (found by
            // compute-sanitizer's slower timing: cudaErrorStreamCaptureInvalidated); an event recorded before that capture began is fine.
            if (!c->pipe_done) CUDA_OK(head, cudaEventCreateWithFlags(&c->pipe_done, cudaEventDisableTiming));
            CUDA_OK(head, cudaEventRecord(c->pipe_done, g.streams[r]));
        }
    }
    for (size_t r = 0; r < n; r++) {
        CUDA_OK(head, cudaSetDevice(g.devices[r]));
        CUDA_OK(head, cudaEventSynchronize(stage_of(head, r)->pipe_done));
    }
    return true;
}

size_t stage_hidden_len(const Model & m, size_t T) { return (size_t) (m.arch_major == 7 ? 2 : 1) * (size_t) m.n_embed * T; }

bool stage_forward(Context * ctx, const uint32_t * tokens, size_t T, const float * hidden_in, float * hidden_out, bool want_logits, cudaStream_t stream) {
    const Model & m = *ctx->model;
    CUDA_OK(ctx, cudaSetDevice(m.dev.device));
    cudaStream_t own = ctx->stream;
    if (stream) ctx->stream = stream;
    ctx->hidden_in = m.layer_begin == 0 ? nullptr : hidden_in;
    ctx->hidden_out = m.layer_end == m.n_layer ? nullptr : hidden_out;
    const bool ok = forward_pass(ctx, tokens, (int) T, want_logits);
    ctx->hidden_in = nullptr; ctx->hidden_out = nullptr;
    ctx->stream = own;
    return ok;
}

}  // namespace rwkv
