"""Reader/writer for the rwkv.cpp ggml model file format -- numpy restatement.

TEST INFRASTRUCTURE (oracle). Follows docs/FILE_FORMAT.md:10-41 and
rwkv_file_format.inc:102-109 (file header), :152-197 (tensor header), rwkv_utilities.inc:1-3 (byte size).
"""
import struct
from collections import OrderedDict

import numpy as np

RWKV_FILE_MAGIC = 0x67676D66  # rwkv.h:23
# rwkv_file_format.inc:5-24 -- on-disk type ids
TYPE_FP32, TYPE_FP16, TYPE_Q4_0, TYPE_Q4_1, TYPE_Q5_0, TYPE_Q5_1, TYPE_Q8_0 = 0, 1, 2, 3, 7, 8, 9
TYPE_NAMES = {0: "FP32", 1: "FP16", 2: "Q4_0", 3: "Q4_1", 7: "Q5_0", 8: "Q5_1", 9: "Q8_0"}
TYPE_IDS = {v: k for k, v in TYPE_NAMES.items()}
# (block elements, block bytes): ggml-common.h:161-221
BLOCK = {TYPE_FP32: (1, 4), TYPE_FP16: (1, 2), TYPE_Q4_0: (32, 18), TYPE_Q4_1: (32, 20),
         TYPE_Q5_0: (32, 22), TYPE_Q5_1: (32, 24), TYPE_Q8_0: (32, 34)}


def tensor_nbytes(dtype, ne):
    """rwkv_utilities.inc:1-3: type_size * ne0*ne1*ne2 / blck_size."""
    blck, size = BLOCK[dtype]
    n = 1
    for d in ne:
        n *= int(d)
    return size * n // blck


class Tensor:
    __slots__ = ("name", "dtype", "ne", "raw")

    def __init__(self, name, dtype, ne, raw):
        self.name, self.dtype, self.ne, self.raw = name, dtype, tuple(int(x) for x in ne), raw

    @property
    def ne3(self):
        return self.ne + (1,) * (3 - len(self.ne))


class ModelFile:
    def __init__(self, version, n_vocab, n_embed, n_layer, data_type, tensors):
        self.version, self.n_vocab, self.n_embed, self.n_layer, self.data_type = version, n_vocab, n_embed, n_layer, data_type
        self.tensors = tensors  # OrderedDict name -> Tensor


def read_model_file(path):
    buf = np.fromfile(path, dtype=np.uint8)
    magic, version, n_vocab, n_embed, n_layer, data_type = struct.unpack_from("<6I", buf, 0)
    assert magic == RWKV_FILE_MAGIC, "bad magic"
    off = 24
    tensors = OrderedDict()
    while off < len(buf):
        dim_count, key_len, dtype = struct.unpack_from("<3I", buf, off)
        off += 12
        ne = struct.unpack_from("<%dI" % dim_count, buf, off)
        off += 4 * dim_count
        name = bytes(buf[off:off + key_len]).decode("utf-8")
        off += key_len
        nbytes = tensor_nbytes(dtype, ne)
        tensors[name] = Tensor(name, dtype, ne, buf[off:off + nbytes])
        off += nbytes
    assert off == len(buf)
    return ModelFile(version, n_vocab, n_embed, n_layer, data_type, tensors)


def write_model_file(path, n_vocab, n_embed, n_layer, data_type, tensors, version=101):
    """tensors: iterable of (name, dtype, ne(tuple, ggml order), raw bytes-like)."""
    with open(path, "wb") as f:
        f.write(struct.pack("<6I", RWKV_FILE_MAGIC, version, n_vocab, n_embed, n_layer, data_type))
        for name, dtype, ne, raw in tensors:
            key = name.encode("utf-8")
            f.write(struct.pack("<3I", len(ne), len(key), dtype))
            f.write(struct.pack("<%dI" % len(ne), *ne))
            f.write(key)
            raw = np.ascontiguousarray(raw).view(np.uint8).reshape(-1)
            assert raw.size == tensor_nbytes(dtype, ne), (name, raw.size, tensor_nbytes(dtype, ne))
            f.write(raw.tobytes())
