// Host-side model quantizer: rwkv_quantize_model_file (reference rwkv_quantize.inc:16-171) with the
// block encoders restated from ggml's deterministic reference quantizers (ggml-quants.c:31-217).
#pragma once
#include <cstddef>
#include <cstdint>
#include "errors.h"

namespace rwkv {

// Encodes n (multiple of 32) floats into blocks of `type`; returns bytes written.
size_t quantize_row(int type, const float * x, void * dst, size_t n);

bool quantize_model_file(const char * in_path, const char * out_path, const char * format_name, ErrorSink sink);

}  // namespace rwkv
