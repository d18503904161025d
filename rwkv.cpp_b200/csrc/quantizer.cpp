#define _FILE_OFFSET_BITS 64
#include "quantizer.h"

#include <cfloat>
#include <cinttypes>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "formats.h"
#include "ggml_file.h"

// The encoders below must round exactly like the reference's scalar code: one rounding per
// operation, no fused multiply-add.
#if defined(__GNUC__)
#pragma GCC optimize("fp-contract=off")
#endif

namespace rwkv {

namespace {

inline void put_half(uint8_t * p, float v) { uint16_t h = fp32_to_fp16(v); memcpy(p, &h, 2); }
inline int min_int(int a, int b) { return a < b ? a : b; }

// quantize_row_q4_0_ref / quantize_row_q5_0_ref (ggml-quants.c:31-66, 107-150): symmetric, the element
// of largest magnitude maps to -8 / -16.
void encode_sym(const float * x, uint8_t * out, int levels /* 8 or 16 */) {
    float amax = 0.0f, max = 0.0f;
    for (int j = 0; j < 32; j++) {
        const float v = x[j];
        if (amax < fabsf(v)) { amax = fabsf(v); max = v; }
    }
    const float d = max / (float) -levels;
    const float id = d ? 1.0f / d : 0.0f;
    put_half(out, d);
    uint8_t * qs = out + (levels == 8 ? 2 : 6);
    uint32_t qh = 0;
    for (int j = 0; j < 16; j++) {
        const float x0 = x[j] * id, x1 = x[16 + j] * id;
        const uint8_t xi0 = (uint8_t) min_int(2 * levels - 1, (int8_t) (x0 + ((float) levels + 0.5f)));
        const uint8_t xi1 = (uint8_t) min_int(2 * levels - 1, (int8_t) (x1 + ((float) levels + 0.5f)));
        qs[j] = (uint8_t) ((xi0 & 0x0F) | ((xi1 & 0x0F) << 4));
        qh |= ((xi0 & 0x10u) >> 4) << (j + 0);
        qh |= ((xi1 & 0x10u) >> 4) << (j + 16);
    }
    if (levels == 16) memcpy(out + 2, &qh, 4);
}

// quantize_row_q4_1_ref / quantize_row_q5_1_ref (ggml-quants.c:68-105, 152-192): min/max affine.
void encode_affine(const float * x, uint8_t * out, int nlev /* 15 or 31 */) {
    float min = FLT_MAX, max = -FLT_MAX;
    for (int j = 0; j < 32; j++) {
        const float v = x[j];
        if (v < min) min = v;
        if (v > max) max = v;
    }
    const float d = (max - min) / (float) nlev;
    const float id = d ? 1.0f / d : 0.0f;
    put_half(out, d);
    put_half(out + 2, min);
    uint8_t * qs = out + (nlev == 15 ? 4 : 8);
    uint32_t qh = 0;
    for (int j = 0; j < 16; j++) {
        const float x0 = (x[j] - min) * id, x1 = (x[16 + j] - min) * id;
        uint8_t xi0, xi1;
        if (nlev == 15) {
            xi0 = (uint8_t) min_int(15, (int8_t) (x0 + 0.5f));
            xi1 = (uint8_t) min_int(15, (int8_t) (x1 + 0.5f));
        } else {
            xi0 = (uint8_t) (x0 + 0.5f);
            xi1 = (uint8_t) (x1 + 0.5f);
        }
        qs[j] = (uint8_t) ((xi0 & 0x0F) | ((xi1 & 0x0F) << 4));
        qh |= ((xi0 & 0x10u) >> 4) << (j + 0);
        qh |= ((xi1 & 0x10u) >> 4) << (j + 16);
    }
    if (nlev == 31) memcpy(out + 4, &qh, 4);
}

// quantize_row_q8_0_ref (ggml-quants.c:194-217)
void encode_q8_0(const float * x, uint8_t * out) {
    float amax = 0.0f;
    for (int j = 0; j < 32; j++) { const float v = fabsf(x[j]); if (v > amax) amax = v; }
    const float d = amax / 127.0f;
    const float id = d ? 1.0f / d : 0.0f;
    put_half(out, d);
    for (int j = 0; j < 32; j++) out[2 + j] = (uint8_t) (int8_t) roundf(x[j] * id);
}

// Tensors the reference never quantizes (rwkv_quantize.inc:1-13).
bool tensor_needs_quant(const std::string & name) {
    static const char * skip[] = {"att.v1", "att.v2", "att.g1", "att.g2", "att.a1", "att.a2", "att.w1", "att.w2", "att.r_k"};
    if (name == "emb.weight" || name == "head.weight") return false;
    for (const char * s : skip) if (name.find(s) != std::string::npos) return false;
    return true;
}

}  // namespace

size_t quantize_row(int type, const float * x, void * dst, size_t n) {
    uint8_t * out = static_cast<uint8_t *>(dst);
    const int bb = dtype_block_bytes(type);
    for (size_t b = 0; b < n / 32; b++) {
        const float * xb = x + b * 32;
        uint8_t * ob = out + b * bb;
        switch (type) {
            case DT_Q4_0: encode_sym(xb, ob, 8); break;
            case DT_Q5_0: encode_sym(xb, ob, 16); break;
            case DT_Q4_1: encode_affine(xb, ob, 15); break;
            case DT_Q5_1: encode_affine(xb, ob, 31); break;
            default: encode_q8_0(xb, ob); break;
        }
    }
    return n / 32 * (size_t) bb;
}

#define MSG(...) do { if (*sink.print) fprintf(stderr, __VA_ARGS__); } while (0)

bool quantize_model_file(const char * in_path, const char * out_path, const char * format_name, ErrorSink sink) {
    RWKV_CHECK(sink, RWKV_ERROR_ARGS, false, in_path && out_path && format_name, "NULL argument");
    const int out_type = dtype_from_name(format_name);
    RWKV_CHECK(sink, RWKV_ERROR_ARGS | RWKV_ERROR_DATA_TYPE, false,
               out_type == DT_Q4_0 || out_type == DT_Q4_1 || out_type == DT_Q5_0 || out_type == DT_Q5_1 || out_type == DT_Q8_0,
               "Unsupported output data type (%s)", format_name);
    MSG("Loading model from '%s'\n", in_path);

    File in(fopen(in_path, "rb"));
    RWKV_CHECK(sink, RWKV_ERROR_FILE | RWKV_ERROR_FILE_OPEN, false, in.f, "Failed to open %s for reading", in_path);
    struct stat st;
    RWKV_CHECK(sink, RWKV_ERROR_FILE | RWKV_ERROR_FILE_STAT, false, fstat(fileno(in.f), &st) == 0, "failed to stat file %s", in_path);
    File out(fopen(out_path, "wb"));
    RWKV_CHECK(sink, RWKV_ERROR_FILE | RWKV_ERROR_FILE_OPEN, false, out.f, "Failed to open %s for writing", out_path);

    FileHeader header;
    {
        bool ok = read_file_header(in.f, header, sink);
        RWKV_CHECK(sink, RWKV_ERROR_FILE, false, ok, "Invalid file header");
    }
    RWKV_CHECK(sink, RWKV_ERROR_FILE, false, header.data_type == DT_F32 || header.data_type == DT_F16,
               "Unsupported input data type (%s); needs to be FP32 or FP16", dtype_name((int) header.data_type));
    FileHeader out_header = header;
    out_header.version = RWKV_FILE_VERSION;
    out_header.data_type = (uint32_t) out_type;
    RWKV_CHECK(sink, RWKV_ERROR_FILE | RWKV_ERROR_FILE_WRITE, false, fwrite(&out_header, sizeof(out_header), 1, out.f) == 1, "Failed to write file header");

    size_t orig_total = 0, new_total = 0;
    std::vector<uint8_t> raw, packed;
    std::vector<float> f32;
    while ((uint64_t) ftello(in.f) < (uint64_t) st.st_size) {
        TensorInfo t;
        {
            bool ok = read_tensor_info(in.f, t, sink);
            RWKV_CHECK(sink, RWKV_ERROR_MODEL_PARAMS, false, ok, "Failed to read tensor header");
        }
        MSG("%48s - [%5" PRIu64 ", %5" PRIu64 ", %5" PRIu64 "], type = %6s ", t.name.c_str(), t.ne[0], t.ne[1], t.ne[2], dtype_name((int) t.data_type));
        raw.resize(t.nbytes);
        RWKV_CHECK(sink, RWKV_ERROR_MODEL_PARAMS | RWKV_ERROR_FILE_READ, false, t.nbytes == 0 || fread(raw.data(), t.nbytes, 1, in.f) == 1,
                   "\nFailed to read tensor data of %s", t.name.c_str());
        const uint8_t * payload = raw.data();
        size_t payload_bytes = t.nbytes;
        uint32_t out_dtype = t.data_type;
        // 2-D FP32/FP16 tensors except embedding/head and the v7 LoRA / r_k tensors (rwkv_quantize.inc:133-140)
        if ((t.data_type == DT_F32 || t.data_type == DT_F16) && t.dim_count == 2 && tensor_needs_quant(t.name)) {
            const size_t n = (size_t) (t.ne[0] * t.ne[1] * t.ne[2]);
            RWKV_CHECK(sink, RWKV_ERROR_MODEL_PARAMS | RWKV_ERROR_SHAPE, false, t.ne[0] % 32 == 0, "\nRow length of %s is not a multiple of 32", t.name.c_str());
            const float * src;
            if (t.data_type == DT_F16) {
                f32.resize(n);
                const uint16_t * h = reinterpret_cast<const uint16_t *>(raw.data());
                for (size_t i = 0; i < n; i++) f32[i] = fp16_to_fp32(h[i]);
                src = f32.data();
            } else {
                src = reinterpret_cast<const float *>(raw.data());
            }
            packed.resize(tensor_nbytes(out_type, t.ne[0], t.ne[1], t.ne[2]));
            payload_bytes = quantize_row(out_type, src, packed.data(), n);
            payload = packed.data();
            out_dtype = (uint32_t) out_type;
            MSG("-> %6s size = %8.2f MB -> %8.2f MB\n", dtype_name(out_type), t.nbytes / 1024.0 / 1024.0, payload_bytes / 1024.0 / 1024.0);
        } else {
            MSG("size = %8.3f MB\n", t.nbytes / 1024.0 / 1024.0);
        }
        // tensor record (rwkv_fwrite_tensor, rwkv_file_format.inc:199-236)
        uint32_t head[6] = {t.dim_count, (uint32_t) t.name.size(), out_dtype, (uint32_t) t.ne[0], (uint32_t) t.ne[1], (uint32_t) t.ne[2]};
        bool ok = fwrite(head, sizeof(uint32_t), 3 + t.dim_count, out.f) == 3 + t.dim_count
               && (t.name.empty() || fwrite(t.name.data(), t.name.size(), 1, out.f) == 1)
               && (payload_bytes == 0 || fwrite(payload, payload_bytes, 1, out.f) == 1);
        RWKV_CHECK(sink, RWKV_ERROR_FILE_WRITE, false, ok, "Failed to write tensor %s", t.name.c_str());
        orig_total += t.nbytes;
        new_total += payload_bytes;
    }
    MSG("original size     = %8.2f MB\n", orig_total / 1024.0 / 1024.0);
    MSG("quantized size    = %8.2f MB\n", new_total / 1024.0 / 1024.0);
    MSG("compression ratio = %8.2f\n", orig_total / (double) new_total);
    return true;
}

}  // namespace rwkv
