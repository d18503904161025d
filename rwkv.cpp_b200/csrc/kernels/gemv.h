// Fused dequantize-and-GEMV: y[M, T] = epilogue(W[M, K] . prologue(x[K, T])) for a batch of
// independent problems in one launch (e.g. r/k/v/g + decay-LoRA of one RWKV layer).
// Replaces ggml_mul_mat on the eval path (reference rwkv_graph.inc:112-116, 244-252, 349-363, 384,
// 504-510, 708; CPU ggml-cpu.c:7377 + quant vec_dot kernels; CUDA ggml-cuda/mmvq.cu:55, mmv.cu:5).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
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
// PRO_LAYERNORM: LN(x; ln_w, ln_b, eps 1e-5) applied while staging x (the head).
enum GemvPrologue : int { PRO_NONE = 0, PRO_LAYERNORM = 1 };

struct GemvProblem {
    const void * W;          // device, rows of quant blocks / f16 / f32, `pitch` bytes apart (16-B multiple)
    const void * Wt;         // device, tile-major prefill copy of W (gemm_tc_repack) or NULL: no tensor-core path
    long long pitch;
    int type, K, M;
    const float * x;  long long ldx;      // input  column t at x + t*ldx
    // Single-token launches of the streaming kernel: x[:, 0] already in the staged layout of this problem's (type, K) (act_stage.cuh),
    // emitted by the kernel that produced x. The consumers copy it instead of quantising x; NULL = stage from x. Ignored by the
    // generic kernel and the tensor-core path.
    const unsigned char * xq;
    float * y;        long long ldy;      // output column t at y + t*ldy
    const float * res;  long long ldres;
    const float * gate; long long ldgate;
    const float * bias;                   // [M]
    const float * ln_w; const float * ln_b;  // [K]
    int epi, pro;
    int first_cta, n_cta;                 // filled by gemv_launch
    int wk, g, tile_rows;                 // streaming kernel only: warps sharing a row along K, lanes per row, rows per TMA tile
};

// In-kernel timeline (our stand-in for nsys, which this image lacks): when a trace slot is attached, thread 0 of every
// CTA folds %globaltimer into [min start, max end] of the slot. Works inside CUDA-graph replays.
struct TraceRec { unsigned long long start, end, mark[4]; };   // mark[]: optional intra-kernel points of CTA 0
// The cursor belongs to the host thread that is enqueuing a pass (contexts are evaluated one thread each, rwkv.h:94-96): thread_local,
// so two clones traced from two threads never share slots; enqueue_pass copies the names into its Context when the pass is complete.
extern thread_local TraceRec * g_trace_base;        // device buffer, or nullptr when tracing is off
extern thread_local int g_trace_next;               // next free slot of the current pass
extern thread_local const char * g_trace_names[1024];
inline TraceRec * trace_slot(const char * name) {
    if (!g_trace_base || g_trace_next >= 1024) return nullptr;
    g_trace_names[g_trace_next] = name;
    return g_trace_base + g_trace_next++;
}

constexpr int GEMV_MAX_PROBLEMS = 8;
struct GemvBatch {
    int n, T;
    TraceRec * trace;
    long long max_col_bytes, stage_bytes; // streaming kernel only: shared-memory carve-up
    GemvProblem p[GEMV_MAX_PROBLEMS];
};

struct DeviceInfo { int device; int num_sms; int max_smem_optin; };

// Enqueues one kernel on `stream` covering all problems. Returns cudaSuccess or the launch error.
// Uses the TMA streaming kernel (gemv_tma.cu) whenever every problem's shape fits it, else the generic
// direct-load kernel (gemv.cu); the choice depends on the problem shapes only, never on T.
cudaError_t gemv_launch(GemvBatch & batch, const DeviceInfo & dev, cudaStream_t stream);
cudaError_t gemv_tma_launch(GemvBatch & batch, const DeviceInfo & dev, cudaStream_t stream);      // cudaErrorNotSupported if a shape does not fit
cudaError_t gemv_generic_launch(GemvBatch & batch, const DeviceInfo & dev, cudaStream_t stream);

// Brings a freshly uploaded matrix (file bytes, rows `pitch` bytes apart) into the device layout of quant_decode.cuh, in place: the
// fifth-bit word of every Q5_0 / Q5_1 block is bit-transposed (qh5_to_device); every other type is left as it is. Must run before any
// kernel reads the matrix (including gemm_tc_repack).
cudaError_t weights_to_device_layout(void * W, long long pitch, int type, int M, int K, cudaStream_t stream);

// Tensor-core (tcgen05) path for chunks of >= 32 tokens (gemm_tc.cu). act16_scratch: device scratch for the fp16 copies of the
// input matrices (sum over distinct inputs of round16(T) * K halves).
bool gemm_tc_supported(const GemvProblem & p, int T);
// The tile-major prefill copy of a matrix the tensor-core path streams (0 bytes = the matrix cannot take that path).
bool gemm_tc_eligible(int type, int K);
size_t gemm_tc_tiled_bytes(int type, int M, int K);
cudaError_t gemm_tc_repack(const void * W, long long pitch, int type, int M, int K, void * dst, cudaStream_t stream);
// workspace: gemm_tc_workspace_bytes(T, sum over the batch's distinct inputs of round16(T) * K) bytes of device memory whose first
// GEMM_TC_COUNTER_BYTES were zeroed once after the allocation (split-K tile counters; every launch leaves them zero again).
constexpr size_t GEMM_TC_COUNTER_BYTES = 4096;
constexpr size_t GEMM_TC_PARTIAL_BYTES = 0;                      // (K-splits are reduced through distributed shared memory: no global partials)
size_t gemm_tc_workspace_bytes(int T, size_t operand_halves);
cudaError_t gemm_tc_launch(GemvBatch & batch, const DeviceInfo & dev, cudaStream_t stream, void * workspace, size_t workspace_bytes);

// Programmatic dependent launch for every kernel of the eval path (RWKV_B200_NO_PDL=1 turns it off).
extern bool g_use_pdl;

// Every kernel of the eval path asks for the SAME L1 / shared-memory split (all shared): an SM can only change its carve-out when it is
// idle, so a stream that alternates between 113-225 KB weight-streaming kernels and small glue kernels with the default (L1-heavy)
// preference makes every kernel boundary wait for the SMs to drain instead of overlapping through programmatic dependent launch.
// Once per (device, kernel); RWKV_B200_CARVEOUT=0 leaves the driver's default (A/B aid).
void prefer_max_shared_carveout(const void * kernel);

// Launch with the PDL attribute: the kernel may start while its predecessor in the stream is still running; it must
// execute griddepcontrol.wait before touching anything the predecessor writes.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args &&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_use_pdl ? 1 : 0;
    prefer_max_shared_carveout(reinterpret_cast<const void *>(kernel));
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#if defined(__CUDACC__)
// First statement of every non-GEMV kernel: let the successor start (it will prefetch its weights), then wait for the
// predecessor's results.
__device__ __forceinline__ void trace_begin(TraceRec * t) {
    if (t && threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); atomicMin(&t->start, g); }
}
__device__ __forceinline__ void trace_end(TraceRec * t) {
    if (t && threadIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); atomicMax(&t->end, g); }
}
__device__ __forceinline__ void trace_mark(TraceRec * t, int i) {
    if (t && threadIdx.x == 0 && blockIdx.x == 0) { unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g)); t->mark[i] = g; }
}
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// Per-launch count of kernels this module has enqueued (bench.py reports it as gpu_launches).
extern std::atomic<unsigned long long> g_kernel_launches;

// One-time per-device function attribute (dynamic shared-memory opt-in): several host threads may launch concurrently.
struct PerDeviceOnce {
    std::atomic<bool> done[64] = {};
    template <typename F> cudaError_t run(F && set) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return set();
        if (done[dev].load(std::memory_order_acquire)) return cudaSuccess;
        const cudaError_t e = set();               // idempotent: a second thread repeating it is harmless
        if (e == cudaSuccess) done[dev].store(true, std::memory_order_release);
        return e;
    }
};

}  // namespace rwkv
