// Per-context execution state and the forward pass (serial + sequence mode) over a resident Model.
// Replaces rwkv_computation_graph + ggml_backend_sched (reference rwkv_graph.inc:16-54, 611-882;
// rwkv_eval.inc:25-35): there is no graph IR -- the launch sequence is written out per architecture.
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdint>
#include <mutex>
#include <vector>

#include "model.h"

namespace rwkv {

// One process, several GPUs, behind the plain rwkv.h API (RWKV_B200_PIPELINE_DEVICES=0,1,...): the layers are cut into stages, one
// per listed device, connected by the peer-memory hand-off (kernels/pipe.cu). The handle the caller holds is the stage-0 context;
// it owns one context per further stage. All contexts of a stage (the original and its clones) enqueue on the stage's ONE stream,
// and one pass over all stages is enqueued under `order`, so items travel through every link in the same order.
struct PipeGroup {
    std::vector<cudaStream_t> streams;     // one per stage, on that stage's device
    std::vector<int> devices;
    std::mutex order;
    std::atomic<int> refs{0};
};

struct CopyPool;                        // helper threads of the pageable-buffer bounce path (engine.cu)
void destroy_copy_pool(CopyPool * p);

struct Context {
    Model * model = nullptr;
    PipeGroup * group = nullptr;           // set on the stage-0 context of an in-process pipeline
    std::vector<Context *> stages;         // ... which owns the contexts of stages 1 .. N-1
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_start = nullptr, ev_stop = nullptr;

    // device buffers
    float * state_a = nullptr;       // current input state
    float * state_b = nullptr;       // output state of the running evaluation
    float * state_init = nullptr;    // rwkv_init_state image, copied when state_in == NULL
    float * logits = nullptr;        // [n_vocab]
    int * tokens = nullptr;          // [capacity_T]
    float * scratch = nullptr;       // activation arena for capacity_T tokens
    unsigned char * xq = nullptr;    // staged activation columns handed from producer kernels to single-token GEMVs (act_stage.cuh)
    size_t xq_slot_bytes = 0;
    cudaEvent_t pipe_done = nullptr; // in-process pipeline: end of this context's latest evaluation on its stage's stream
    void * act16 = nullptr;          // fp16 copies of GEMM inputs for the tensor-core prefill path
    size_t act16_bytes = 0;
    bool use_tensor_cores = true;
    size_t scratch_floats = 0;
    int capacity_T = 0;
    // Two pinned token slots, used alternately (`phase`): the host may prepare pass n+1 while pass n runs.
    int * tokens_host[2] = {nullptr, nullptr};
    cudaEvent_t slot_free[2] = {nullptr, nullptr};   // recorded after the pass that last read the slot
    bool slot_used[2] = {false, false};
    int phase = 0;                   // flips every pass together with the state_a/state_b swap

    // A single-token pass runs either over all resident layers at once (segment 0) or, when the caller's host state is copied in
    // and out around it, as n_segments consecutive layer groups (segments 1..n_segments) so that the H2D copy of group g+1 and the
    // D2H copy of group g-1 overlap the kernels of group g. Slots below are keyed by slot_index(want_logits, phase, segment).
    static constexpr int MAX_SEGMENTS = 8;
    static constexpr int N_SLOTS = 2 * 2 * (1 + MAX_SEGMENTS);
    static int slot_index(bool want_logits, int phase, int segment) { return ((want_logits ? 1 : 0) * 2 + phase) * (1 + MAX_SEGMENTS) + segment; }
    int n_segments = 1;              // layer groups of the overlapped path (1 = no overlap)
    bool overlap_copies = false;     // rwkv_eval with host state: pipeline the state copies against the layer groups
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t seg_in[MAX_SEGMENTS] = {}, seg_out[MAX_SEGMENTS] = {};
    cudaEvent_t pass_begin = nullptr;
    // pinned bounce buffers for pageable caller memory (engine.cu: eval_host_overlapped), allocated on first need
    float * bounce_in = nullptr, * bounce_out = nullptr, * bounce_logits = nullptr;
    cudaEvent_t seg_d2h[MAX_SEGMENTS] = {};          // slice g has arrived in bounce_out
    CopyPool * copy_pool = nullptr;

    // CUDA graphs for single-token passes; captured on the second use of a slot.
    struct GraphSlot { cudaGraphExec_t exec = nullptr; int uses = 0; unsigned long long launches = 0; };
    GraphSlot graphs[N_SLOTS];
    bool use_graphs = true;

    // Profiling mode (bench.py roofline leg): CUDA events around every GEMV launch, graphs off.
    bool profiling = false;
    struct ProfRecord { cudaEvent_t start, stop; double bytes; };
    std::vector<ProfRecord> prof;

    TraceRec * trace_buf = nullptr;  // in-kernel timeline records (rwkv_b200_trace_*), 1024 slots
    int trace_count = 0;             // slots used by the last enqueued / captured pass
    const char * trace_names[1024] = {};   // kernel name per slot (string literals)

    // Pipeline-stage hand-off: device pointers of the caller for the pass being enqueued, or NULL.
    // Layout: x f32[C x T], then (v7 only) v_first f32[C x T].
    const float * hidden_in = nullptr;
    float * hidden_out = nullptr;

    // Batched multi-sequence decode (kernels/batch.cu, rwkv_b200_batch_*): batch_n > 0 makes this a batch context whose state
    // buffers hold batch_n states back to back (batch_stride = state_len floats apart) and whose logits buffer holds batch_n columns;
    // every pass evaluates exactly one token of each sequence.
    int batch_n = 0;
    long long batch_stride = 0;

    // On-device sampling (kernels/sampling.cu): result word, scratch for biased logits, device copy of the caller's logit bias.
    bool logits_valid = false;       // ctx->logits holds the head output of the most recent pass
    uint32_t * sample_token = nullptr;   // device
    uint32_t * sample_token_host = nullptr;   // pinned
    float * sample_scratch = nullptr;    // [n_vocab] device
    uint32_t * bias_ids = nullptr; float * bias_values = nullptr; size_t bias_capacity = 0;   // device

    float last_device_ms = 0.f;      // CUDA-event time of the last forward (kernels only)
    int last_error = 0;              // rwkv_error_flags
    bool print_errors = true;
    uint32_t n_threads = 1;

    ErrorSink sink() { return ErrorSink{&last_error, &print_errors}; }
};

// Largest number of tokens pushed through the kernels in one go; longer sequences are cut into
// pieces of this size with the state staying on the device.
constexpr int MAX_TOKENS_PER_PASS = 256;
extern std::atomic<size_t> g_bounce_min_bytes;   // pageable caller states of at least this many bytes travel through pinned bounce buffers

Context * create_context(Model * model, ErrorSink sink, int batch_n = 0);   // batch_n > 0: a batch context for that many sequences
void destroy_context(Context * ctx);

// In-process pipeline over `devices` (one stage per entry, layer blocks balanced by bytes with the head on the last stage).
Context * create_pipeline(const char * path, const std::vector<int> & devices, ErrorSink sink);
Context * clone_pipeline(Context * head, ErrorSink sink);
// rwkv_eval / rwkv_eval_sequence / rwkv_eval_sequence_in_chunks for a pipeline handle: every stage takes its slice of the caller's
// host state, passes of <= MAX_TOKENS_PER_PASS tokens run through all stages (the state stays resident in between), every stage
// hands its slice back, the last one the logits. chunk == 0: plain sequence evaluation.
bool pipeline_eval_host(Context * head, const uint32_t * tokens, size_t T, size_t chunk, const float * state_in, float * state_out, float * logits_out);

// Host image of a fresh state (rwkv_init_state, rwkv_eval.inc:224-241).
void fill_init_state(const Model & m, float * state);

// state_a <- host state (or the init image when NULL). Asynchronous on ctx->stream.
bool upload_state(Context * ctx, const float * state_in);
// host <- state_a / logits; synchronises the stream.
bool download_outputs(Context * ctx, float * state_out, float * logits_out);

// rwkv_eval / rwkv_eval_sequence with caller-owned host buffers: state_in (NULL = fresh state) -> T tokens -> state_out / logits_out
// (either may be NULL). With ctx->overlap_copies and a pass that fits one launch sequence the state travels per layer group,
// overlapped with the kernels; otherwise upload_state + forward + download_outputs. Synchronises before returning.
bool eval_host(Context * ctx, const uint32_t * tokens, size_t T, const float * state_in, float * state_out, float * logits_out);

// rwkv_eval_sequence_in_chunks: the reference's chunk loop (rwkv_eval.inc:158-222) with the state resident between chunks.
bool eval_host_chunks(Context * ctx, const uint32_t * tokens, size_t T, size_t chunk, const float * state_in, float * state_out, float * logits_out);

// Runs T tokens (host pointer) through all resident layers: reads state_a, leaves the new state in
// state_a (buffers are swapped internally), and, if want_logits, ln_out + head of the last token
// into ctx->logits. Asynchronous except for the token upload.
bool forward(Context * ctx, const uint32_t * tokens, size_t T, bool want_logits);

// One pipeline stage: the resident layers [layer_begin, layer_end) for T <= MAX_TOKENS_PER_PASS tokens. The first stage
// embeds `tokens`, the others start from `hidden_in`; every stage but the last leaves its result in `hidden_out`
// (device pointers, stage_hidden_len(T) floats); the last one computes the logits when asked. `stream` (may be NULL = the
// context's own) is the CUDA stream everything is enqueued on, so the hand-off can be ordered against NCCL sends and
// receives without host synchronisation.
// Batch contexts: state of sequence `seq` <- host state (NULL = fresh) / -> host; one token of every sequence; logits of `seq`.
bool batch_set_state(Context * ctx, int seq, const float * state_in);
bool batch_get_state(Context * ctx, int seq, float * state_out);
bool batch_eval(Context * ctx, const uint32_t * tokens, bool want_logits);
bool batch_get_logits(Context * ctx, int seq, float * logits_out);

// Samples the next token from ctx->logits on the device (reference python/sampling.py:10-52 with the caller's uniform number u in
// [0, 1) in place of numpy's RandomState draw); only the token id is copied back. Requires a preceding pass that computed logits.
bool sample_token(Context * ctx, float temperature, float top_p, double u, const uint32_t * bias_ids, const float * bias_values, size_t n_bias, uint32_t * token_out);

size_t stage_hidden_len(const Model & m, size_t T);

bool stage_forward(Context * ctx, const uint32_t * tokens, size_t T, const float * hidden_in, float * hidden_out, bool want_logits, cudaStream_t stream);
// The same with the hand-off done by peer-memory kernels on the stream (kernels/pipe.cu): receive from the previous stage's
// stores (unless this stage embeds tokens), run the resident layers, store into the next stage's mailbox (unless this is the last).
bool pipe_forward(Context * ctx, const uint32_t * tokens, size_t T, bool want_logits, cudaStream_t stream);
// Allocates (once) the stage's mailbox for passes of up to MAX_TOKENS_PER_PASS tokens.
bool pipe_ensure_box(Context * ctx);

}  // namespace rwkv
