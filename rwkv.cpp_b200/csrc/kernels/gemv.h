// Fused dequantize-and-GEMV: y[M, T] = epilogue(W[M, K] . prologue(x[K, T])) for a batch of
// independent problems in one launch (e.g. r/k/v/g + decay-LoRA of one RWKV layer).
// Replaces ggml_mul_mat on the eval path (reference rwkv_graph.inc:112-116, 244-252, 349-363, 384,
// 504-510, 708; CPU ggml-cpu.c:7377 + quant vec_dot kernels; CUDA ggml-cuda/mmvq.cu:55, mmv.cu:5).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace rwkv {

enum GemvEpilogue : int {
    EPI_NONE = 0,
    EPI_SIGMOID,          // 1/(1+exp(-v))
    EPI_SILU,             // v/(1+exp(-v))
    EPI_TANH,
    EPI_RELU_SQR,         // max(v,0)^2
    EPI_ADD,              // res + v                       (residual add after att.output)
    EPI_MUL_ADD,          // res + gate*v                  (x + r*(Wv k) of the channel mix)
    EPI_BIAS_EXPNEGEXP,   // exp(-exp(v + bias[row]))      (v6 decay, rwkv_graph.inc:365-367)
    EPI_BIAS_SIGMOID,     // sigmoid(v + bias[row])        (v7 a / v-gate, rwkv_graph.inc:417-423, 445-450)
    EPI_BIAS_W7,          // exp(-0.606531*sigmoid(v+bias[row]))  (v7 decay, rwkv_graph.inc:425-430)
};
enum GemvPrologue : int { PRO_NONE = 0, PRO_LAYERNORM = 1 };   // LN(x; ln_w, ln_b, eps 1e-5) applied while staging x

struct GemvProblem {
    const void * W;          // device, rows of quant blocks / f16 / f32, `pitch` bytes apart (16-B multiple)
    long long pitch;
    int type, K, M;
    const float * x;  long long ldx;      // input  column t at x + t*ldx
    float * y;        long long ldy;      // output column t at y + t*ldy
    const float * res;  long long ldres;
    const float * gate; long long ldgate;
    const float * bias;                   // [M]
    const float * ln_w; const float * ln_b;  // [K]
    int epi, pro;
    int first_cta, n_cta;                 // filled by gemv_launch
};

constexpr int GEMV_MAX_PROBLEMS = 8;
struct GemvBatch {
    int n, T;
    GemvProblem p[GEMV_MAX_PROBLEMS];
};

struct DeviceInfo { int device; int num_sms; int max_smem_optin; };

// Enqueues one kernel on `stream` covering all problems. Returns cudaSuccess or the launch error.
cudaError_t gemv_launch(GemvBatch & batch, const DeviceInfo & dev, cudaStream_t stream);

// Per-launch count of kernels this module has enqueued (bench.py reports it as gpu_launches).
extern unsigned long long g_kernel_launches;

}  // namespace rwkv
