// The rwkv.h C ABI over the B200 engine, plus the additive rwkv_b200_* entry points.
// Argument checks, error categories and NULL conventions follow the reference entry points
// (rwkv.cpp:71-258, rwkv_eval.inc:38-241) so its C tests run unchanged against this library.
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/rwkv.h"
#include "../../include/rwkv_b200.h"
#include "engine.h"
#include "kernels/ops.h"
#include "quantizer.h"

using namespace rwkv;

// The opaque handle of the C API is the engine's Context.
struct rwkv_context : public rwkv::Context {};

static inline Context * C(struct rwkv_context * c) { return static_cast<Context *>(c); }
static inline const Context * C(const struct rwkv_context * c) { return static_cast<const Context *>(c); }

static int default_device() {
    const char * e = getenv("RWKV_B200_DEVICE");
    return e ? atoi(e) : 0;
}

extern "C" {

struct rwkv_context * rwkv_b200_init_from_file_ex(const char * file_path, int device, int layer_begin, int layer_end) {
    g_last_error = RWKV_ERROR_NONE;
    ErrorSink sink = global_sink();
    RWKV_CHECK(sink, RWKV_ERROR_ARGS, nullptr, file_path != nullptr, "Model file path is NULL");
    Model * model = load_model(file_path, device, layer_begin, layer_end, sink);
    RWKV_PROPAGATE(sink, nullptr, model != nullptr);
    Context * ctx = create_context(model, sink);   // deletes the model itself on failure
    RWKV_PROPAGATE(sink, nullptr, ctx != nullptr);
    return static_cast<struct rwkv_context *>(ctx);
}

// RWKV_B200_PIPELINE_DEVICES="0,1,2,3": one stage per listed CUDA device (a device may repeat), all inside this process
static std::vector<int> pipeline_devices() {
    std::vector<int> d;
    const char * e = getenv("RWKV_B200_PIPELINE_DEVICES");
    if (!e) return d;
    for (const char * p = e; *p;) {
        char * end = nullptr;
        const long v = strtol(p, &end, 10);
        if (end == p) break;
        d.push_back((int) v);
        p = *end == ',' ? end + 1 : end;
    }
    return d;
}

struct rwkv_context * rwkv_b200_init_pipeline(const char * file_path, const int * devices, size_t n_devices) {
    g_last_error = RWKV_ERROR_NONE;
    ErrorSink sink = global_sink();
    RWKV_CHECK(sink, RWKV_ERROR_ARGS, nullptr, file_path && devices && n_devices >= 2, "A pipeline needs a model file and at least two devices");
    Context * head = create_pipeline(file_path, std::vector<int>(devices, devices + n_devices), sink);
    RWKV_PROPAGATE(sink, nullptr, head != nullptr);
    return static_cast<struct rwkv_context *>(head);
}

size_t rwkv_b200_pipeline_stages(const struct rwkv_context * ctx) { return ctx && C(ctx)->group ? C(ctx)->stages.size() + 1 : 0; }

static bool whole_model(const Context * c) {
    return (c->group || (c->model->layer_begin == 0 && c->model->layer_end == c->model->n_layer)) && c->batch_n == 0;
}

struct rwkv_context * rwkv_init_from_file(const char * file_path, const uint32_t n_threads, const uint32_t n_gpu_layers) {
    (void) n_gpu_layers;   // every layer always lives on the GPU
    const std::vector<int> devs = pipeline_devices();
    struct rwkv_context * ctx = devs.size() >= 2 ? rwkv_b200_init_pipeline(file_path, devs.data(), devs.size())
                                                 : rwkv_b200_init_from_file_ex(file_path, default_device(), 0, -1);
    if (ctx) C(ctx)->n_threads = n_threads;
    return ctx;
}

struct rwkv_context * rwkv_clone_context(struct rwkv_context * ctx, const uint32_t n_threads) {
    if (!ctx) return nullptr;
    bool print = C(ctx)->print_errors;
    int flags = 0;
    ErrorSink sink{&flags, &print};
    Context * clone = C(ctx)->group ? clone_pipeline(C(ctx), sink) : create_context(C(ctx)->model, sink);
    if (!clone) { g_last_error |= flags; return nullptr; }
    clone->n_threads = n_threads;
    clone->print_errors = C(ctx)->print_errors;
    return static_cast<struct rwkv_context *>(clone);
}

bool rwkv_eval(struct rwkv_context * ctx, const uint32_t token, const float * state_in, float * state_out, float * logits_out) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    const size_t n_vocab = (size_t) c->model->n_vocab;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, token < n_vocab, "Token (%" PRIu32 ") is out of range (0 .. %zu)", token, n_vocab - 1);
    RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, whole_model(c),
               "This context holds only a pipeline stage or is a batch context; use the rwkv_b200 stage / batch API");
    if (c->group) return pipeline_eval_host(c, &token, 1, 0, state_in, state_out, logits_out);
    return eval_host(c, &token, 1, state_in, state_out, logits_out);
}

bool rwkv_eval_sequence(struct rwkv_context * ctx, const uint32_t * sequence, const size_t sequence_len, const float * state_in, float * state_out, float * logits_out) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, sequence_len > 0, "Sequence length is 0");
    if (!sequence) return true;   // "build the graph only" in the reference (rwkv_eval.inc:122); nothing to build here
    const size_t n_vocab = (size_t) c->model->n_vocab;
    for (size_t i = 0; i < sequence_len; i++)
        RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, sequence[i] < n_vocab, "Token at index %zu (%" PRIu32 ") is out of range (0 .. %zu)", i, sequence[i], n_vocab - 1);
    RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, whole_model(c),
               "This context holds only a pipeline stage or is a batch context; use the rwkv_b200 stage / batch API");
    if (c->group) return pipeline_eval_host(c, sequence, sequence_len, 0, state_in, state_out, logits_out);
    return eval_host(c, sequence, sequence_len, state_in, state_out, logits_out);
}

bool rwkv_eval_sequence_in_chunks(struct rwkv_context * ctx, const uint32_t * tokens, const size_t sequence_len, const size_t chunk_size,
                                  const float * state_in, float * state_out, float * logits_out) {
    Context * c = C(ctx);
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, sequence_len > 0, "Sequence length is 0");
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, chunk_size > 0, "Chunk size is 0");
    c->last_error = RWKV_ERROR_NONE;
    if (!tokens) return true;
    const size_t n_vocab = (size_t) c->model->n_vocab;
    for (size_t i = 0; i < sequence_len; i++)
        RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, tokens[i] < n_vocab, "Token at index %zu (%" PRIu32 ") is out of range (0 .. %zu)", i, tokens[i], n_vocab - 1);
    RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, whole_model(c),
               "This context holds only a pipeline stage or is a batch context; use the rwkv_b200 stage / batch API");
    if (c->group) return pipeline_eval_host(c, tokens, sequence_len, chunk_size, state_in, state_out, logits_out);
    return eval_host_chunks(c, tokens, sequence_len, chunk_size, state_in, state_out, logits_out);
}

size_t rwkv_get_n_vocab(const struct rwkv_context * ctx) { return (size_t) C(ctx)->model->n_vocab; }
size_t rwkv_get_n_embed(const struct rwkv_context * ctx) { return (size_t) C(ctx)->model->n_embed; }
size_t rwkv_get_n_layer(const struct rwkv_context * ctx) { return (size_t) C(ctx)->model->n_layer; }
size_t rwkv_get_state_len(const struct rwkv_context * ctx) { return C(ctx)->model->state_len(); }
size_t rwkv_get_logits_len(const struct rwkv_context * ctx) { return (size_t) C(ctx)->model->n_vocab; }
uint32_t rwkv_get_state_buffer_element_count(const struct rwkv_context * ctx) { return (uint32_t) rwkv_get_state_len(ctx); }
uint32_t rwkv_get_logits_buffer_element_count(const struct rwkv_context * ctx) { return (uint32_t) rwkv_get_logits_len(ctx); }

void rwkv_init_state(const struct rwkv_context * ctx, float * state) { fill_init_state(*C(ctx)->model, state); }

void rwkv_free(struct rwkv_context * ctx) {
    if (!ctx) return;
    destroy_context(C(ctx));
}

void rwkv_set_print_errors(struct rwkv_context * ctx, const bool print_errors) {
    if (ctx) C(ctx)->print_errors = print_errors; else g_print_errors = print_errors;
}
bool rwkv_get_print_errors(const struct rwkv_context * ctx) { return ctx ? C(ctx)->print_errors : g_print_errors; }
enum rwkv_error_flags rwkv_get_last_error(struct rwkv_context * ctx) {
    int * p = ctx ? &C(ctx)->last_error : &g_last_error;
    int v = *p;
    *p = RWKV_ERROR_NONE;
    return (enum rwkv_error_flags) v;
}

bool rwkv_quantize_model_file(const char * in_path, const char * out_path, const char * format_name) {
    g_last_error = RWKV_ERROR_NONE;
    return quantize_model_file(in_path, out_path, format_name, global_sink());
}

const char * rwkv_get_system_info_string(void) {
    static std::string s;
    if (s.empty()) {
        s = "RWKV_B200=1 ARCH=sm_100a";
        int n = 0;
        if (cudaGetDeviceCount(&n) == cudaSuccess && n > 0) {
            cudaDeviceProp p;
            int drv = 0, rt = 0;
            cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt);
            if (cudaGetDeviceProperties(&p, default_device() < n ? default_device() : 0) == cudaSuccess) {
                s += " CUDA_DEVICES=" + std::to_string(n) + " DEVICE=\"" + p.name + "\" SM=" + std::to_string(p.major) + std::to_string(p.minor) +
                     " SMS=" + std::to_string(p.multiProcessorCount) + " HBM_GB=" + std::to_string((unsigned long long) (p.totalGlobalMem >> 30)) +
                     " DRIVER=" + std::to_string(drv) + " RUNTIME=" + std::to_string(rt);
            }
        } else {
            s += " CUDA_DEVICES=0";
        }
    }
    return s.c_str();
}

// ---- additive entry points (include/rwkv_b200.h) ----------------------------------------------------

bool rwkv_b200_inspect_file(const char * path, struct rwkv_b200_file_info * out) {
    g_last_error = RWKV_ERROR_NONE;
    ErrorSink sink = global_sink();
    RWKV_CHECK(sink, RWKV_ERROR_ARGS, false, path && out, "NULL argument");
    ModelFile mf;
    RWKV_PROPAGATE(sink, false, scan_model_file(path, mf, sink));
    memset(out, 0, sizeof(*out));
    out->version = mf.header.version; out->n_vocab = mf.header.n_vocab; out->n_embed = mf.header.n_embed;
    out->n_layer = mf.header.n_layer; out->data_type = mf.header.data_type;
    out->n_tensors = mf.tensors.size(); out->file_size = mf.file_size;
    out->arch_major = 4; out->arch_minor = 0;
    if (mf.find("blocks.0.att.ln_x.weight")) { out->arch_major = 5; out->arch_minor = mf.find("blocks.0.att.gate.weight") ? 2 : 1; }
    if (mf.find("blocks.0.att.time_maa_x")) { out->arch_major = 6; out->arch_minor = 0; }
    if (mf.find("blocks.0.att.r_k")) { out->arch_major = 7; out->arch_minor = 0; }
    if (out->arch_major == 7) out->head_count = (uint32_t) mf.find("blocks.0.att.r_k")->ne[1];
    else if (out->arch_major >= 5) {
        const TensorInfo * td = mf.find("blocks.0.att.time_decay");
        RWKV_CHECK(sink, RWKV_ERROR_MODEL_PARAMS | RWKV_ERROR_PARAM_MISSING, false, td, "Model parameter blocks.0.att.time_decay not found");
        out->head_count = (uint32_t) td->ne[2];
    }
    if (out->head_count) out->head_size = out->n_embed / out->head_count;
    out->state_len = (uint64_t) out->n_embed * (out->arch_major >= 5 ? 2 + out->head_size : 5) * out->n_layer;
    uint64_t bytes = 0;
    for (const TensorInfo & t : mf.tensors) bytes += (t.name == "emb.weight") ? t.nbytes / t.ne[1] : t.nbytes;
    out->bytes_per_token = bytes + 2 * 4 * out->state_len;
    return true;
}

bool rwkv_b200_state_load(struct rwkv_context * ctx, const float * state_in) { C(ctx)->last_error = 0; return C(ctx)->batch_n == 0 && upload_state(C(ctx), state_in); }
bool rwkv_b200_state_store(struct rwkv_context * ctx, float * state_out) { return C(ctx)->batch_n == 0 && download_outputs(C(ctx), state_out, nullptr); }
bool rwkv_b200_synchronize(struct rwkv_context * ctx) { return download_outputs(C(ctx), nullptr, nullptr); }

bool rwkv_b200_eval_resident(struct rwkv_context * ctx, const uint32_t * tokens, size_t n_tokens, bool want_logits, float * logits_out) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, tokens && n_tokens > 0, "No tokens");
    RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, c->batch_n == 0, "This is a batch context; use rwkv_b200_batch_eval");
    const size_t n_vocab = (size_t) c->model->n_vocab;
    for (size_t i = 0; i < n_tokens; i++)
        RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, tokens[i] < n_vocab, "Token at index %zu (%" PRIu32 ") is out of range (0 .. %zu)", i, tokens[i], n_vocab - 1);
    if (!forward(c, tokens, n_tokens, want_logits || logits_out)) return false;
    if (logits_out) return download_outputs(c, nullptr, logits_out);
    return true;
}

struct rwkv_context * rwkv_b200_batch_create(struct rwkv_context * ctx, size_t n_sequences) {
    if (!ctx) return nullptr;
    Context * c = C(ctx);
    bool print = c->print_errors;
    int flags = 0;
    ErrorSink sink{&flags, &print};
    const Model & m = *c->model;
    Context * b = nullptr;
    if (n_sequences == 0 || n_sequences > (size_t) MAX_TOKENS_PER_PASS) {
        flags |= RWKV_ERROR_ARGS;
        if (print) fprintf(stderr, "A batch holds 1 .. %d sequences\n", MAX_TOKENS_PER_PASS);
    } else if (m.layer_begin != 0 || m.layer_end != m.n_layer || !batch_shape_supported(m.arch_major, m.n_embed, m.head_size)) {
        flags |= RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED;
        if (print) fprintf(stderr, "Batched decode supports whole RWKV v4 / v5 / v6 / v7 models with n_embed <= 4096 and head size <= 64 (v7: 128)\n");
    } else {
        b = create_context(c->model, sink, (int) n_sequences);
    }
    if (!b) { g_last_error |= flags; return nullptr; }
    b->print_errors = c->print_errors;
    return static_cast<struct rwkv_context *>(b);
}
bool rwkv_b200_batch_set_state(struct rwkv_context * batch, size_t sequence, const float * state_in) {
    C(batch)->last_error = RWKV_ERROR_NONE;
    return batch_set_state(C(batch), (int) sequence, state_in);
}
bool rwkv_b200_batch_get_state(struct rwkv_context * batch, size_t sequence, float * state_out) {
    C(batch)->last_error = RWKV_ERROR_NONE;
    return batch_get_state(C(batch), (int) sequence, state_out);
}
bool rwkv_b200_batch_eval(struct rwkv_context * batch, const uint32_t * tokens, bool want_logits) {
    Context * c = C(batch);
    c->last_error = RWKV_ERROR_NONE;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, c->batch_n > 0 && tokens, "Not a batch context or NULL tokens");
    for (int i = 0; i < c->batch_n; i++)
        RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, tokens[i] < (uint32_t) c->model->n_vocab, "Token of sequence %d (%" PRIu32 ") is out of range (0 .. %d)", i, tokens[i], c->model->n_vocab - 1);
    return batch_eval(c, tokens, want_logits);
}
bool rwkv_b200_batch_get_logits(struct rwkv_context * batch, size_t sequence, float * logits_out) {
    C(batch)->last_error = RWKV_ERROR_NONE;
    return batch_get_logits(C(batch), (int) sequence, logits_out);
}
size_t rwkv_b200_batch_size(const struct rwkv_context * ctx) { return (size_t) C(ctx)->batch_n; }

bool rwkv_b200_sample(struct rwkv_context * ctx, float temperature, float top_p, double u, const uint32_t * bias_ids, const float * bias_values, size_t n_bias,
                      uint32_t * token_out) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    return sample_token(c, temperature, top_p, u, bias_ids, bias_values, n_bias, token_out);
}

bool rwkv_b200_sample_logits(const float * logits, size_t n_vocab, float temperature, float top_p, double u, const uint32_t * bias_ids, const float * bias_values,
                             size_t n_bias, uint32_t * token_out, float * prob_out) {
    g_last_error = RWKV_ERROR_NONE;
    ErrorSink sink = global_sink();
    RWKV_CHECK(sink, RWKV_ERROR_ARGS, false, logits && token_out && n_vocab > 0 && n_vocab <= 65536 && temperature >= 0.0f && top_p >= 0.0f && top_p <= 1.0f &&
               u >= 0.0 && u < 1.0 && (n_bias == 0 || (bias_ids && bias_values)), "Invalid sampling arguments");
    int dev = default_device(), n_dev = 0;
    RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, cudaGetDeviceCount(&n_dev) == cudaSuccess && dev < n_dev && cudaSetDevice(dev) == cudaSuccess,
               "No usable CUDA device; this engine has no CPU execution path");
    float * d_logits = nullptr, * d_scratch = nullptr, * d_bv = nullptr, * d_prob = nullptr;
    uint32_t * d_bi = nullptr, * d_tok = nullptr;
    bool ok = cudaMalloc((void **) &d_logits, n_vocab * 4) == cudaSuccess && cudaMalloc((void **) &d_scratch, n_vocab * 4) == cudaSuccess &&
              cudaMalloc((void **) &d_tok, 4) == cudaSuccess && cudaMalloc((void **) &d_prob, 4) == cudaSuccess &&
              cudaMemcpy(d_logits, logits, n_vocab * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && n_bias) ok = cudaMalloc((void **) &d_bi, n_bias * 4) == cudaSuccess && cudaMalloc((void **) &d_bv, n_bias * 4) == cudaSuccess &&
                           cudaMemcpy(d_bi, bias_ids, n_bias * 4, cudaMemcpyHostToDevice) == cudaSuccess && cudaMemcpy(d_bv, bias_values, n_bias * 4, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok) {
        SampleParams sp{};
        sp.logits = d_logits; sp.n_vocab = (int) n_vocab; sp.temperature = temperature; sp.top_p = top_p; sp.u = u;
        sp.bias_ids = d_bi; sp.bias_values = d_bv; sp.n_bias = (int) n_bias; sp.scratch = d_scratch; sp.token_out = d_tok; sp.prob_out = d_prob;
        ok = launch_sample(sp, 0) == cudaSuccess && cudaDeviceSynchronize() == cudaSuccess && cudaMemcpy(token_out, d_tok, 4, cudaMemcpyDeviceToHost) == cudaSuccess;
        if (ok && prob_out) ok = cudaMemcpy(prob_out, d_prob, 4, cudaMemcpyDeviceToHost) == cudaSuccess;
    }
    cudaError_t e = cudaGetLastError();
    cudaFree(d_logits); cudaFree(d_scratch); cudaFree(d_tok); cudaFree(d_prob); cudaFree(d_bi); cudaFree(d_bv);
    RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, ok, "sampling failed on the device: %s", cudaGetErrorString(e));
    return true;
}

bool rwkv_b200_eval_sample(struct rwkv_context * ctx, uint32_t token, float temperature, float top_p, double u, uint32_t * next_token_out) {
    if (!rwkv_b200_eval_resident(ctx, &token, 1, true, nullptr)) return false;
    return sample_token(C(ctx), temperature, top_p, u, nullptr, nullptr, 0, next_token_out);
}

size_t rwkv_b200_stage_hidden_len(const struct rwkv_context * ctx, size_t n_tokens) { return stage_hidden_len(*C(ctx)->model, n_tokens); }

bool rwkv_b200_stage_eval(struct rwkv_context * ctx, const uint32_t * tokens, size_t n_tokens, const float * hidden_in, float * hidden_out, bool want_logits,
                          void * cuda_stream) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    const Model & m = *c->model;
    const bool first = m.layer_begin == 0, last = m.layer_end == m.n_layer;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, n_tokens > 0 && n_tokens <= (size_t) MAX_TOKENS_PER_PASS, "A stage pass takes 1 .. %d tokens", MAX_TOKENS_PER_PASS);
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, first ? tokens != nullptr : hidden_in != nullptr, "Stage [%d, %d) needs %s", m.layer_begin, m.layer_end,
               first ? "tokens" : "the previous stage's activations");
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, last || hidden_out != nullptr, "Stage [%d, %d) needs an output buffer for its activations", m.layer_begin, m.layer_end);
    if (first)
        for (size_t i = 0; i < n_tokens; i++)
            RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, tokens[i] < (uint32_t) m.n_vocab, "Token at index %zu (%" PRIu32 ") is out of range (0 .. %d)", i, tokens[i], m.n_vocab - 1);
    return stage_forward(c, tokens, n_tokens, hidden_in, hidden_out, want_logits && last, reinterpret_cast<cudaStream_t>(cuda_stream));
}

// ---- layer pipeline over peer memory (kernels/pipe.cu) ----
void * rwkv_b200_stream(struct rwkv_context * ctx) { return ctx ? reinterpret_cast<void *>(C(ctx)->stream) : nullptr; }
size_t rwkv_b200_pipe_handle_size(void) { return sizeof(cudaIpcMemHandle_t); }

bool rwkv_b200_pipe_export(struct rwkv_context * ctx, void * handle_out) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, handle_out != nullptr, "handle_out is NULL");
    if (!pipe_ensure_box(c)) return false;
    cudaIpcMemHandle_t h;
    const cudaError_t e = cudaIpcGetMemHandle(&h, c->model->link.box);
    RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, e == cudaSuccess, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    memcpy(handle_out, &h, sizeof(h));
    return true;
}

static bool pipe_open(Context * c, const void * handle, PipeBox ** out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void * p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, e == cudaSuccess, "cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
    *out = reinterpret_cast<PipeBox *>(p);
    return true;
}

bool rwkv_b200_pipe_connect(struct rwkv_context * ctx, const void * prev_handle, const void * next_handle) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    Model & m = *c->model;
    if (!pipe_ensure_box(c)) return false;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, (m.layer_begin == 0) == (prev_handle == nullptr) && (m.layer_end == m.n_layer) == (next_handle == nullptr),
               "Stage [%d, %d): a first stage has no previous neighbour, a last stage no next one, every other stage both", m.layer_begin, m.layer_end);
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, !m.link.prev && !m.link.next, "The stage is already connected");
    if (cudaSetDevice(m.dev.device) != cudaSuccess) return false;
    if (prev_handle) { if (!pipe_open(c, prev_handle, &m.link.prev)) return false; m.link.prev_ipc = true; }
    if (next_handle) { if (!pipe_open(c, next_handle, &m.link.next)) return false; m.link.next_ipc = true; }
    return true;
}

bool rwkv_b200_pipe_connect_local(struct rwkv_context * ctx, struct rwkv_context * prev, struct rwkv_context * next) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    Model & m = *c->model;
    if (!pipe_ensure_box(c)) return false;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, (m.layer_begin == 0) == (prev == nullptr) && (m.layer_end == m.n_layer) == (next == nullptr),
               "Stage [%d, %d): a first stage has no previous neighbour, a last stage no next one, every other stage both", m.layer_begin, m.layer_end);
    for (struct rwkv_context * other : {prev, next}) {
        if (!other) continue;
        Context * o = C(other);
        if (!pipe_ensure_box(o)) return false;
        if (o->model->dev.device != m.dev.device) {      // one process, two GPUs: peer access in both directions
            for (int pass = 0; pass < 2; pass++) {
                const int from = pass ? o->model->dev.device : m.dev.device, to = pass ? m.dev.device : o->model->dev.device;
                int can = 0;
                cudaDeviceCanAccessPeer(&can, from, to);
                RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, can, "Device %d cannot access device %d's memory", from, to);
                cudaSetDevice(from);
                const cudaError_t e = cudaDeviceEnablePeerAccess(to, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, e == cudaSuccess, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e));
            }
        }
        (other == prev ? m.link.prev : m.link.next) = o->model->link.box;
    }
    return true;
}

bool rwkv_b200_pipe_eval(struct rwkv_context * ctx, const uint32_t * tokens, size_t n_tokens, bool want_logits, void * cuda_stream) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    const Model & m = *c->model;
    const bool first = m.layer_begin == 0;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, n_tokens > 0 && n_tokens <= (size_t) MAX_TOKENS_PER_PASS, "A stage pass takes 1 .. %d tokens", MAX_TOKENS_PER_PASS);
    RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, c->batch_n == 0, "Batch contexts cannot be pipeline stages");
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, !first || tokens != nullptr, "The first stage needs tokens");
    if (first)
        for (size_t i = 0; i < n_tokens; i++)
            RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, tokens[i] < (uint32_t) m.n_vocab, "Token at index %zu (%" PRIu32 ") is out of range (0 .. %d)", i, tokens[i], m.n_vocab - 1);
    return pipe_forward(c, tokens, n_tokens, want_logits, reinterpret_cast<cudaStream_t>(cuda_stream));
}

bool rwkv_b200_stage_logits(struct rwkv_context * ctx, float * logits_out, void * cuda_stream) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, logits_out && c->model->layer_end == c->model->n_layer, "Only the last stage holds logits");
    cudaStream_t s = cuda_stream ? reinterpret_cast<cudaStream_t>(cuda_stream) : c->stream;
    cudaError_t e = cudaMemcpyAsync(logits_out, c->logits, (size_t) c->model->n_vocab * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    RWKV_CHECK(c->sink(), RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, e == cudaSuccess, "Reading the logits failed: %s", cudaGetErrorString(e));
    return true;
}

float rwkv_b200_last_device_ms(const struct rwkv_context * ctx) {
    float ms = -1.f;
    const Context * c = C(ctx);
    if (cudaEventSynchronize(c->ev_stop) != cudaSuccess) return -1.f;
    if (cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop) != cudaSuccess) return -1.f;
    return ms;
}

uint64_t rwkv_b200_kernel_launch_count(void) { return g_kernel_launches; }

uint64_t rwkv_b200_gemv_bytes_per_token(const struct rwkv_context * ctx, bool with_logits) {
    const Model & m = *C(ctx)->model;
    return (uint64_t) (m.gemv_bytes_per_token - (with_logits ? 0 : m.head_matrix_bytes));
}

uint64_t rwkv_b200_bytes_per_token(const struct rwkv_context * ctx, bool with_logits) {
    const Model & m = *C(ctx)->model;
    return (uint64_t) (m.weight_bytes_per_token - (with_logits ? 0 : m.head_bytes) + 2 * 4 * m.state_len());
}

float rwkv_b200_time_resident(struct rwkv_context * ctx, const uint32_t * tokens, size_t tokens_per_step, int n_steps, int warmup_steps, bool want_logits) {
    Context * c = C(ctx);
    cudaEvent_t e0, e1;
    if (!tokens || tokens_per_step == 0 || n_steps <= 0 || cudaSetDevice(c->model->dev.device) != cudaSuccess) return -1.f;
    const uint32_t * t = tokens;
    for (int i = 0; i < warmup_steps; i++, t += tokens_per_step) if (!rwkv_b200_eval_resident(ctx, t, tokens_per_step, want_logits, nullptr)) return -1.f;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return -1.f;
    cudaStreamSynchronize(c->stream);
    cudaEventRecord(e0, c->stream);
    bool ok = true;
    for (int i = 0; i < n_steps && ok; i++, t += tokens_per_step) ok = rwkv_b200_eval_resident(ctx, t, tokens_per_step, want_logits, nullptr);
    cudaEventRecord(e1, c->stream);
    float ms = -1.f;
    if (ok && cudaEventSynchronize(e1) == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return ok ? ms : -1.f;
}

bool rwkv_b200_profile_pass(struct rwkv_context * ctx, const uint32_t * tokens, size_t n_tokens, bool want_logits, struct rwkv_b200_profile * out) {
    Context * c = C(ctx);
    c->last_error = RWKV_ERROR_NONE;
    RWKV_CHECK(c->sink(), RWKV_ERROR_ARGS, false, tokens && out && n_tokens > 0 && n_tokens <= (size_t) MAX_TOKENS_PER_PASS, "Invalid profile arguments");
    memset(out, 0, sizeof(*out));
    c->profiling = true;
    const unsigned long long before = g_kernel_launches;
    bool ok = rwkv_b200_eval_resident(ctx, tokens, n_tokens, want_logits, nullptr) && cudaStreamSynchronize(c->stream) == cudaSuccess;
    c->profiling = false;
    out->total_launches = (uint32_t) (g_kernel_launches - before);
    if (ok) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev_start, c->ev_stop);
        out->pass_ms = ms;
        for (auto & r : c->prof) {
            cudaEventElapsedTime(&ms, r.start, r.stop);
            out->gemv_ms += ms; out->gemv_bytes += r.bytes; out->gemv_launches++;
            if (ms > out->top_ms) { out->top_ms = ms; out->top_bytes = r.bytes; }
        }
    }
    for (auto & r : c->prof) { cudaEventDestroy(r.start); cudaEventDestroy(r.stop); }
    c->prof.clear();
    return ok;
}

void rwkv_b200_set_graphs(struct rwkv_context * ctx, bool enabled) { C(ctx)->use_graphs = enabled; }
void rwkv_b200_set_tensor_cores(struct rwkv_context * ctx, bool enabled) { C(ctx)->use_tensor_cores = enabled; }
void rwkv_b200_set_overlap(struct rwkv_context * ctx, bool enabled) { C(ctx)->overlap_copies = enabled; }
void rwkv_b200_set_bounce_min_bytes(size_t bytes) { g_bounce_min_bytes.store(bytes); }
int rwkv_b200_overlap_groups(const struct rwkv_context * ctx) { return C(ctx)->overlap_copies && C(ctx)->n_segments > 1 ? C(ctx)->n_segments : 0; }
static bool trace_rearm(Context * c) {
    std::vector<TraceRec> init(1024);
    for (auto & r : init) { r.start = ~0ull; r.end = 0; for (auto & m : r.mark) m = 0; }
    return cudaMemcpy(c->trace_buf, init.data(), init.size() * sizeof(TraceRec), cudaMemcpyHostToDevice) == cudaSuccess;
}

bool rwkv_b200_trace_enable(struct rwkv_context * ctx) {
    Context * c = C(ctx);
    if (cudaSetDevice(c->model->dev.device) != cudaSuccess || cudaStreamSynchronize(c->stream) != cudaSuccess) return false;
    if (!c->trace_buf && cudaMalloc((void **) &c->trace_buf, 1024 * sizeof(TraceRec)) != cudaSuccess) return false;
    // graphs captured so far carry no trace slots: drop them so they are re-captured
    for (auto & g : c->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); g = Context::GraphSlot(); }
    return trace_rearm(c);
}

void rwkv_b200_trace_disable(struct rwkv_context * ctx) {
    Context * c = C(ctx);
    if (!c->trace_buf || cudaSetDevice(c->model->dev.device) != cudaSuccess) return;
    cudaStreamSynchronize(c->stream);
    for (auto & g : c->graphs) { if (g.exec) cudaGraphExecDestroy(g.exec); g = Context::GraphSlot(); }
    cudaFree(c->trace_buf);
    c->trace_buf = nullptr; c->trace_count = 0;
}

// marks_us: optional [max_records][4] intra-kernel marks of CTA 0 (-1 when absent)
static double * g_trace_marks_out = nullptr;
void rwkv_b200_trace_set_marks_buffer(double * marks_us) { g_trace_marks_out = marks_us; }

int rwkv_b200_trace_read(struct rwkv_context * ctx, double * start_us, double * end_us, char (*names)[32], int max_records) {
    Context * c = C(ctx);
    if (!c->trace_buf || cudaStreamSynchronize(c->stream) != cudaSuccess) return -1;
    std::vector<TraceRec> recs(1024);
    if (cudaMemcpy(recs.data(), c->trace_buf, recs.size() * sizeof(TraceRec), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    const int n = c->trace_count < max_records ? c->trace_count : max_records;
    unsigned long long t0 = ~0ull;
    for (int i = 0; i < n; i++) if (recs[i].start < t0) t0 = recs[i].start;
    for (int i = 0; i < n; i++) {
        start_us[i] = recs[i].start == ~0ull ? -1.0 : (double) (recs[i].start - t0) * 1e-3;
        end_us[i] = recs[i].end == 0 ? -1.0 : (double) (recs[i].end - t0) * 1e-3;
        if (g_trace_marks_out) for (int k = 0; k < 4; k++) g_trace_marks_out[i * 4 + k] = recs[i].mark[k] ? (double) (recs[i].mark[k] - t0) * 1e-3 : -1.0;
        const char * nm = c->trace_names[i] ? c->trace_names[i] : "?";
        strncpy(names[i], nm, 31); names[i][31] = 0;
    }
    trace_rearm(c);
    return n;
}

bool rwkv_b200_matvec(int data_type, int K, int M, int T, const void * weights, const float * x, float * y, int epilogue) {
    g_last_error = RWKV_ERROR_NONE;
    ErrorSink sink = global_sink();
    RWKV_CHECK(sink, RWKV_ERROR_ARGS, false, weights && x && y && K > 0 && M > 0 && T > 0 && dtype_supported(data_type) && K % dtype_block_elems(data_type) == 0 && epilogue >= 0 && epilogue <= EPI_RELU_SQR,
               "Invalid matvec arguments");
    int dev = default_device(), n_dev = 0;
    RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, cudaGetDeviceCount(&n_dev) == cudaSuccess && dev < n_dev && cudaSetDevice(dev) == cudaSuccess,
               "No usable CUDA device; this engine has no CPU execution path");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    DeviceInfo di{dev, prop.multiProcessorCount, (int) prop.sharedMemPerBlockOptin};
    const size_t row_bytes = tensor_nbytes(data_type, (uint64_t) K, 1, 1);
    size_t unit_bytes = row_bytes;
    if (data_type == DT_Q4_0 || data_type == DT_Q5_0 || data_type == DT_Q8_0) unit_bytes = (size_t) ((K / 32 + 1) / 2 * 2) * dtype_block_bytes(data_type);
    const size_t pitch = (unit_bytes + 15) / 16 * 16;
    uint8_t * dW = nullptr; float * dx = nullptr, * dy = nullptr;
    bool ok = cudaMalloc((void **) &dW, pitch * M + 256) == cudaSuccess && cudaMalloc((void **) &dx, sizeof(float) * K * T) == cudaSuccess &&
              cudaMalloc((void **) &dy, sizeof(float) * M * T) == cudaSuccess;
    ok = ok && cudaMemset(dW, 0, pitch * M + 256) == cudaSuccess &&
         cudaMemcpy2D(dW, pitch, weights, row_bytes, row_bytes, M, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(dx, x, sizeof(float) * K * T, cudaMemcpyHostToDevice) == cudaSuccess &&
         weights_to_device_layout(dW, (long long) pitch, data_type, M, K, 0) == cudaSuccess;       // what the model loader does after an upload
    if (ok) {
        GemvBatch b;
        memset(&b, 0, sizeof(b));
        b.n = 1; b.T = T;
        GemvProblem & p = b.p[0];
        p.W = dW; p.pitch = (long long) pitch; p.type = data_type; p.K = K; p.M = M;
        p.x = dx; p.ldx = K; p.y = dy; p.ldy = M; p.epi = epilogue; p.pro = PRO_NONE;
        void * act16 = nullptr, * tiled = nullptr;
        const size_t act16_bytes = gemm_tc_workspace_bytes(T, (size_t) ((T + 15) / 16 * 16) * K + 128);
        // passes of >= 32 tokens go through the tensor-core kernel, which streams the tile-major copy of the matrix
        if (T >= 32 && gemm_tc_eligible(data_type, K) && !getenv("RWKV_B200_NO_TC")) {
            ok = cudaMalloc(&tiled, gemm_tc_tiled_bytes(data_type, M, K)) == cudaSuccess && gemm_tc_repack(dW, (long long) pitch, data_type, M, K, tiled, 0) == cudaSuccess;
            p.Wt = tiled;
        }
        const bool use_tc = ok && gemm_tc_supported(p, T);
        if (use_tc) ok = cudaMalloc(&act16, act16_bytes) == cudaSuccess && cudaMemset(act16, 0, GEMM_TC_COUNTER_BYTES) == cudaSuccess && gemm_tc_launch(b, di, 0, act16, act16_bytes) == cudaSuccess;
        else ok = gemv_launch(b, di, 0) == cudaSuccess;
        ok = ok && cudaDeviceSynchronize() == cudaSuccess && cudaMemcpy(y, dy, sizeof(float) * M * T, cudaMemcpyDeviceToHost) == cudaSuccess;
        cudaFree(act16); cudaFree(tiled);
    }
    cudaError_t e = cudaGetLastError();
    cudaFree(dW); cudaFree(dx); cudaFree(dy);
    RWKV_CHECK(sink, RWKV_ERROR_CTX | RWKV_ERROR_UNSUPPORTED, false, ok, "matvec failed on the device: %s", cudaGetErrorString(e));
    return true;
}

}  // extern "C"
