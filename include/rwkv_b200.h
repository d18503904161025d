/* rwkv_b200.h -- additive entry points of the B200 engine that the reference ABI has no slot for.
 * Nothing here is needed by a drop-in consumer; bench.py, the multi-GPU layer pipeline and the
 * tests use them. All follow the conventions of rwkv.h (bool / NULL returns, flags readable with
 * rwkv_get_last_error).
 */
#ifndef RWKV_B200_H
#define RWKV_B200_H

#include "rwkv.h"

#if defined(__cplusplus)
extern "C" {
#endif

/* Host-only summary of a model file (no GPU needed): what rwkv_init_from_file would load.
 * Architecture detection as reference rwkv_model_loading.inc:319-340, head geometry :403-409. */
struct rwkv_b200_file_info {
    uint32_t version, n_vocab, n_embed, n_layer, data_type;
    uint32_t arch_major, arch_minor, head_count, head_size;
    uint64_t n_tensors, file_size;
    uint64_t state_len;          /* floats, rwkv.cpp:171-179 */
    uint64_t bytes_per_token;    /* SURVEY.md 8(d) byte model: all layer tensors + head + 1 emb row + 2x state */
};
RWKV_API bool rwkv_b200_inspect_file(const char * model_file_path, struct rwkv_b200_file_info * out);

/* rwkv_init_from_file on an explicit CUDA device, keeping only layers [layer_begin, layer_end)
 * resident (layer_end < 0 = n_layer). A context that lacks layer 0 takes its input activations
 * from the previous pipeline stage, one that lacks the last layer cannot produce logits. */
RWKV_API struct rwkv_context * rwkv_b200_init_from_file_ex(const char * model_file_path, int device, int layer_begin, int layer_end);

/* Device-resident evaluation: the recurrent state stays in HBM between calls.
 *   rwkv_b200_state_load : state_in (host, may be NULL = fresh) -> device
 *   rwkv_b200_eval_resident: run n tokens from the resident state; logits stay on the device unless logits_out != NULL
 *   rwkv_b200_state_store: device -> state_out (host) */
RWKV_API bool rwkv_b200_state_load(struct rwkv_context * ctx, const float * state_in);
RWKV_API bool rwkv_b200_eval_resident(struct rwkv_context * ctx, const uint32_t * tokens, size_t n_tokens, bool want_logits, float * logits_out);
RWKV_API bool rwkv_b200_state_store(struct rwkv_context * ctx, float * state_out);
/* Blocks until everything enqueued on the context's stream has finished. */
RWKV_API bool rwkv_b200_synchronize(struct rwkv_context * ctx);

/* Batched multi-sequence decode (the reference has no equivalent -- it evaluates one context per call): a batch
 * context holds the recurrent states of n_sequences independent sequences in HBM and evaluates ONE token of EACH per call, so every
 * weight matrix is streamed once for n_sequences tokens (the fused dequantize-GEMV runs its multi-column path). Per sequence the
 * result is bit-identical to evaluating that sequence alone with rwkv_eval (below 32 sequences; from 32 on the tensor-core path
 * takes the layer matrices, with its own rounding of the activations). RWKV v4 / v5 / v6, n_embed <= 4096, head size <= 64.
 *   rwkv_b200_batch_create    : from any context of the model (shares its weights); free with rwkv_free; NULL on failure
 *   rwkv_b200_batch_set_state : state of one sequence <- state_in (host or device pointer, NULL = fresh state)
 *   rwkv_b200_batch_eval      : tokens[n_sequences], one per sequence; want_logits = run the head for every sequence
 *   rwkv_b200_batch_get_logits / rwkv_b200_batch_get_state : copy one sequence's logits / state out (host or device pointer)
 * rwkv_eval* and the resident / stage entry points refuse a batch context. */
RWKV_API struct rwkv_context * rwkv_b200_batch_create(struct rwkv_context * ctx, size_t n_sequences);
RWKV_API bool rwkv_b200_batch_set_state(struct rwkv_context * batch, size_t sequence, const float * state_in);
RWKV_API bool rwkv_b200_batch_get_state(struct rwkv_context * batch, size_t sequence, float * state_out);
RWKV_API bool rwkv_b200_batch_eval(struct rwkv_context * batch, const uint32_t * tokens, bool want_logits);
RWKV_API bool rwkv_b200_batch_get_logits(struct rwkv_context * batch, size_t sequence, float * logits_out);
RWKV_API size_t rwkv_b200_batch_size(const struct rwkv_context * ctx);   /* 0 for an ordinary context */

/* On-device sampling: the next token is drawn on the GPU from the logits of the most recent evaluation that
 * computed them, following the reference's python/sampling.py:10-52 (softmax, optional logit bias, temperature 0 = argmax, top-p
 * cutoff, power by 1/temperature, renormalise, inverse-CDF draw); only the 4-byte token id crosses PCIe instead of n_vocab floats.
 * `u` is a uniform number in [0, 1) from the caller's generator -- with u = numpy's RandomState.random_sample() the result is the
 * token numpy.random.choice would return in the reference. temperature >= 0, 0 <= top_p <= 1 (0 means 1), n_vocab <= 65536.
 * rwkv_b200_eval_sample = rwkv_b200_eval_resident(token, with logits) + rwkv_b200_sample without bias. */
RWKV_API bool rwkv_b200_sample(struct rwkv_context * ctx, float temperature, float top_p, double u, const uint32_t * bias_ids, const float * bias_values,
                               size_t n_bias, uint32_t * token_out);
RWKV_API bool rwkv_b200_eval_sample(struct rwkv_context * ctx, uint32_t token, float temperature, float top_p, double u, uint32_t * next_token_out);

/* Test hook: the same sampling kernel on caller-provided host logits (any n_vocab <= 65536), no model needed.
 * prob_out (optional) receives the probability the chosen token had after top-p / temperature. */
RWKV_API bool rwkv_b200_sample_logits(const float * logits, size_t n_vocab, float temperature, float top_p, double u, const uint32_t * bias_ids,
                                      const float * bias_values, size_t n_bias, uint32_t * token_out, float * prob_out);

/* Layer pipeline across GPUs (the reference has no equivalent: ggml offloads layers of ONE context,
 * rwkv.cpp:97-116). A context created with rwkv_b200_init_from_file_ex(path, device, begin, end) is one stage; its slice
 * of the recurrent state stays resident on that device. Per pass of n_tokens (<= 256):
 *   first stage : embeds `tokens`, writes its output activations to hidden_out
 *   inner stage : reads hidden_in, writes hidden_out
 *   last stage  : reads hidden_in, computes the logits when want_logits (fetch them with rwkv_b200_stage_logits)
 * hidden_in / hidden_out are DEVICE pointers to rwkv_b200_stage_hidden_len(ctx, n_tokens) floats: x f32[n_embed x n_tokens]
 * followed, for RWKV v7, by v_first f32[n_embed x n_tokens]; the caller moves them between GPUs (NCCL send/recv or a peer
 * copy). cuda_stream (a cudaStream_t, NULL = the context's own stream) is where the pass is enqueued, so it can be ordered
 * against the transfers without host synchronisation (note that the legacy default stream's handle IS NULL: create a stream).
 * A single-stage context (all layers) accepts the call too. */
RWKV_API size_t rwkv_b200_stage_hidden_len(const struct rwkv_context * ctx, size_t n_tokens);
RWKV_API bool rwkv_b200_stage_eval(struct rwkv_context * ctx, const uint32_t * tokens, size_t n_tokens, const float * hidden_in, float * hidden_out,
                                   bool want_logits, void * cuda_stream);
RWKV_API bool rwkv_b200_stage_logits(struct rwkv_context * ctx, float * logits_out, void * cuda_stream);

/* The same pipeline with the hand-off INSIDE the library: every stage owns a mailbox in its HBM that the previous stage fills with
 * NVLink peer stores from the last kernel of its pass, and a credit word the next stage writes back (csrc/kernels/pipe.cu) -- no
 * NCCL call, no host synchronisation and no data-dependent launch parameter per token, so single-token passes keep replaying
 * their CUDA graph. Set-up (once): every stage exports a handle, the handles travel between the ranks by any means (an
 * all_gather of rwkv_b200_pipe_handle_size() bytes; NCCL / gloo are needed for nothing else), every stage connects to its
 * neighbours. One process driving several GPUs uses rwkv_b200_pipe_connect_local instead (peer access, no IPC).
 *   rwkv_b200_pipe_eval: receive (unless first stage) -> resident layers -> send (unless last stage), all enqueued on cuda_stream
 *   (NULL = the context's own). Every context of a stage shares the stage's link: enqueue a link's passes in item order on ONE
 *   stream. The last stage computes logits when asked (rwkv_b200_stage_logits fetches them). A neighbour that never shows up
 *   makes the waiting kernel trap after ~20 s instead of hanging the GPU. */
/* The whole pipeline inside ONE process, behind the plain rwkv.h entry points: rwkv_b200_init_pipeline loads one stage per listed CUDA
 * device (layer blocks balanced by bytes, the head on the last one), connects them with peer access, and returns a handle that
 * rwkv_eval / rwkv_eval_sequence / rwkv_eval_sequence_in_chunks / rwkv_clone_context / rwkv_free accept like any other: every stage
 * takes and returns its own slice of the caller's state buffer, the token ids enter stage 0, the logits leave the last stage.
 * rwkv_init_from_file does the same when the environment holds RWKV_B200_PIPELINE_DEVICES="0,1,2,3" -- an unchanged consumer of the
 * reference binding reaches several GPUs that way (capacity; clones evaluated from several threads fill the stages concurrently). */
RWKV_API struct rwkv_context * rwkv_b200_init_pipeline(const char * model_file_path, const int * devices, size_t n_devices);
RWKV_API size_t rwkv_b200_pipeline_stages(const struct rwkv_context * ctx);   /* 0 for an ordinary context */
RWKV_API void * rwkv_b200_stream(struct rwkv_context * ctx);            /* the context's own cudaStream_t */
RWKV_API size_t rwkv_b200_pipe_handle_size(void);
RWKV_API bool rwkv_b200_pipe_export(struct rwkv_context * ctx, void * handle_out);
RWKV_API bool rwkv_b200_pipe_connect(struct rwkv_context * ctx, const void * prev_handle, const void * next_handle);
RWKV_API bool rwkv_b200_pipe_connect_local(struct rwkv_context * ctx, struct rwkv_context * prev, struct rwkv_context * next);
RWKV_API bool rwkv_b200_pipe_eval(struct rwkv_context * ctx, const uint32_t * tokens, size_t n_tokens, bool want_logits, void * cuda_stream);

/* Measurement hooks. */
RWKV_API float rwkv_b200_last_device_ms(const struct rwkv_context * ctx);     /* CUDA-event time of the last pass */
RWKV_API uint64_t rwkv_b200_kernel_launch_count(void);                         /* kernels enqueued by this process */
RWKV_API uint64_t rwkv_b200_bytes_per_token(const struct rwkv_context * ctx, bool with_logits);
/* The part of that byte model the fused dequantize-GEMV kernel streams: every layer matrix (+ the head), at file precision. */
RWKV_API uint64_t rwkv_b200_gemv_bytes_per_token(const struct rwkv_context * ctx, bool with_logits);
/* Times `n_steps` back-to-back resident passes of `tokens_per_step` tokens each (after `warmup_steps` untimed ones)
 * with CUDA events on the context's own stream. `tokens` holds (warmup_steps + n_steps) * tokens_per_step ids.
 * Returns the milliseconds of the timed steps, or a negative value on error. Single-token passes replay a CUDA graph. */
RWKV_API float rwkv_b200_time_resident(struct rwkv_context * ctx, const uint32_t * tokens, size_t tokens_per_step, int n_steps, int warmup_steps, bool want_logits);

/* Roofline leg: runs ONE resident pass of n_tokens with CUDA events around every fused dequantize-GEMV launch (graphs off)
 * and reports their summed duration, the algorithmic weight bytes they streamed (rwkv_tensor_nbytes of each matrix, reference
 * rwkv_utilities.inc:1-3), their count, and the duration of the whole pass. */
struct rwkv_b200_profile {
    double gemv_ms, gemv_bytes, pass_ms;
    double top_ms, top_bytes;      /* the single slowest GEMV launch of the pass */
    uint32_t gemv_launches, total_launches;
};
RWKV_API bool rwkv_b200_profile_pass(struct rwkv_context * ctx, const uint32_t * tokens, size_t n_tokens, bool want_logits, struct rwkv_b200_profile * out);

/* In-kernel timeline (stand-in for nsys): after rwkv_b200_trace_enable every kernel of a pass folds %globaltimer of its first
 * and last CTA into one record, also inside CUDA-graph replays. rwkv_b200_trace_read returns the records of the most recent
 * pass (microseconds relative to the first kernel's start, names up to 31 chars) and re-arms the buffer. */
RWKV_API bool rwkv_b200_trace_enable(struct rwkv_context * ctx);
RWKV_API void rwkv_b200_trace_disable(struct rwkv_context * ctx);      /* frees the buffer; graphs are re-captured without trace slots */
RWKV_API void rwkv_b200_trace_set_marks_buffer(double * marks_us);   /* optional [max_records][4]: intra-kernel marks of CTA 0 */
RWKV_API int rwkv_b200_trace_read(struct rwkv_context * ctx, double * start_us, double * end_us, char (*names)[32], int max_records);

/* Enables / disables CUDA-graph replay of single-token passes (on by default). */
RWKV_API void rwkv_b200_set_graphs(struct rwkv_context * ctx, bool enabled);
/* Enables / disables the tcgen05 tensor-core kernel for passes of >= 32 tokens (on by default; off = batch-invariant SIMT path). */
RWKV_API void rwkv_b200_set_tensor_cores(struct rwkv_context * ctx, bool enabled);

/* rwkv_eval / rwkv_eval_sequence with caller-owned HOST state: copy the state per layer group (n_layer / 4 groups, at most 8) on
 * dedicated copy streams so that the host-to-device copy of group g+1 and the device-to-host copy of group g-1 overlap the kernels
 * of group g (the ABI state layout is layer-major, so a group is one contiguous slice). Results are bit-identical to the plain
 * upload - evaluate - download order. On by default; RWKV_B200_OVERLAP=0/1 sets the default of new contexts. rwkv_eval_sequence_in_chunks
 * pipelines the upload against its first chunk and the download against its last one. */
RWKV_API void rwkv_b200_set_overlap(struct rwkv_context * ctx, bool enabled);
/* Pageable caller buffers (what the reference's bindings pass: python/rwkv_cpp/rwkv_cpp_model.py:330-351): with overlap on, a host
 * state of at least `bytes` bytes (default 4 MiB; process-wide) that is neither pinned nor device memory is copied slice by slice
 * through pinned bounce buffers of the context by two helper threads, in step with the layer groups, instead of by the driver's
 * synchronous pageable staging. Same bytes either way. RWKV_B200_NO_BOUNCE=1 turns the bounce path off. */
RWKV_API void rwkv_b200_set_bounce_min_bytes(size_t bytes);
/* Number of layer groups the overlapped path uses for this context (0 = overlap off or a single group). */
RWKV_API int rwkv_b200_overlap_groups(const struct rwkv_context * ctx);

/* Test hook: one fused dequantize-GEMV on host buffers, y[M,T] = W[M,K] . x[K,T] (column-major activations),
 * through exactly the kernel the eval path uses (csrc/kernels/gemv.cu). `weights` holds M rows in the file
 * layout of `data_type` (rwkv_file_format.inc:5-24 ids; ggml quant blocks / f16 / f32, unpadded).
 * epilogue: 0 none, 1 sigmoid, 2 silu, 3 tanh, 4 relu^2. */
RWKV_API bool rwkv_b200_matvec(int data_type, int K, int M, int T, const void * weights, const float * x, float * y, int epilogue);

#if defined(__cplusplus)
}
#endif
#endif /* RWKV_B200_H */
