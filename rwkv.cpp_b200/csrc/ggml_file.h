// Host-side reader for the rwkv.cpp ggml model container (docs/FILE_FORMAT.md:10-41;
// reference rwkv_file_format.inc:102-316). No CUDA here: usable and tested without a GPU.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <unordered_map>
#include <vector>
#include "errors.h"
#include "formats.h"

namespace rwkv {

struct FileHeader {            // rwkv_file_format.inc:102-109
    uint32_t magic, version, n_vocab, n_embed, n_layer, data_type;
};

struct TensorInfo {            // rwkv_file_format.inc:152-161 + where the payload sits in the file
    std::string name;
    uint32_t dim_count = 0;
    uint32_t data_type = 0;
    uint64_t ne[3] = {1, 1, 1};   // ggml order: ne[0] is the contiguous (input) dimension
    uint64_t file_offset = 0;     // byte offset of the payload
    size_t nbytes = 0;
};

struct ModelFile {
    FileHeader header{};
    std::vector<TensorInfo> tensors;                       // file order
    std::unordered_map<std::string, size_t> index;         // name -> position in `tensors`
    uint64_t file_size = 0;
    const TensorInfo * find(const std::string & name) const {
        auto it = index.find(name);
        return it == index.end() ? nullptr : &tensors[it->second];
    }
};

// RAII FILE* (reference rwkv_model_loading.inc:114-124).
struct File {
    FILE * f = nullptr;
    explicit File(FILE * f) : f(f) {}
    ~File() { if (f) fclose(f); }
    File(const File &) = delete;
    File & operator=(const File &) = delete;
};

// Validates the 24-byte header exactly as rwkv_fread_file_header (rwkv_file_format.inc:115-142).
bool read_file_header(FILE * f, FileHeader & header, ErrorSink sink);
// Reads one tensor header + name, leaves the stream at the payload (rwkv_file_format.inc:167-197, 240-275).
bool read_tensor_info(FILE * f, TensorInfo & info, ErrorSink sink);
// Opens `path`, reads the header and walks every tensor record to EOF without loading payloads
// (pass 1 of rwkv_load_model_from_file, rwkv_model_loading.inc:288-317).
bool scan_model_file(const char * path, ModelFile & out, ErrorSink sink);

}  // namespace rwkv
