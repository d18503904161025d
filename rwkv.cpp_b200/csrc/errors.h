// Error conventions of the rwkv.h C ABI (reference rwkv_error_handling.inc:1-95):
//   - load / quantize errors OR their flags into a thread-local word, eval errors into the context;
//   - when printing is enabled the message and "file:line: expr" go to stderr;
//   - the failing entry point returns false / NULL; nothing ever aborts the process.
#pragma once
#include <cstdio>
#include "../../include/rwkv.h"

namespace rwkv {

extern thread_local int g_last_error;     // rwkv_error_flags bits
extern thread_local bool g_print_errors;

struct ErrorSink {                        // where a failing check records itself
    int * flags;
    const bool * print;
};
inline ErrorSink global_sink() { return ErrorSink{&g_last_error, &g_print_errors}; }

}  // namespace rwkv

#define RWKV_FAIL_IMPL(sink, err, ret, cond_str, ...)                                   \
    do {                                                                                \
        rwkv::ErrorSink _s = (sink);                                                    \
        *_s.flags |= (int) (err);                                                       \
        if (*_s.print) {                                                                \
            fprintf(stderr, __VA_ARGS__);                                               \
            fprintf(stderr, "\n%s:%d: %s\n", __FILE__, __LINE__, cond_str);             \
        }                                                                               \
        return ret;                                                                     \
    } while (0)

// CHECK(sink, error-flags, return-value, condition, printf-style message...)
#define RWKV_CHECK(sink, err, ret, cond, ...)                                           \
    do { if (!(cond)) RWKV_FAIL_IMPL(sink, err, ret, #cond, __VA_ARGS__); } while (0)

// Propagate a failure that already recorded its flags (reference RWKV_ENSURE_OR_*).
#define RWKV_PROPAGATE(sink, ret, cond)                                                 \
    do { if (!(cond)) { rwkv::ErrorSink _s = (sink);                                    \
         if (*_s.print) fprintf(stderr, "\n%s:%d: %s\n", __FILE__, __LINE__, #cond);    \
         return ret; } } while (0)
