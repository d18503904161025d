"""ctypes loader for the UNMODIFIED reference library built by oracle/Makefile into oracle/_ref/.

TEST INFRASTRUCTURE ONLY. May be imported from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs -- never from the product package (rwkv.cpp_b200/).

The binding follows the reference's own Python binding (python/rwkv_cpp/rwkv_cpp_shared_library.py:49-107)
so the same wrapper can drive either library.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

P_FLOAT = ctypes.POINTER(ctypes.c_float)
P_U32 = ctypes.POINTER(ctypes.c_uint32)


def _host_cpu_flags():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return set(line.split(":", 1)[1].split())
    except OSError:
        pass
    return set()


def reference_library_path():
    """Pick the -march=native build when this host has every CPU flag of the build host, else x86-64-v3."""
    native = os.path.join(REF_DIR, "librwkv_ref_native.so")
    flags_file = os.path.join(REF_DIR, "native.cpuflags")
    portable = os.path.join(REF_DIR, "librwkv_ref.so")
    if os.path.isfile(native) and os.path.isfile(flags_file):
        need = set(open(flags_file).read().split())
        # only ISA extensions matter for SIGILL safety
        isa = {f for f in need if f.startswith(("avx", "sse", "ssse", "fma", "f16c", "bmi", "amx", "vnni", "popcnt", "movbe", "lzcnt", "gfni", "vaes", "vpclmul", "sha", "adx", "rdseed", "rdrnd", "clwb", "clflushopt"))}
        if isa <= _host_cpu_flags():
            return native
    if os.path.isfile(portable):
        return portable
    return None


def bind_rwkv_api(lib):
    """Declare argtypes/restype for the rwkv.h C API (rwkv.h:76-221, rwkv.cpp:145,151)."""
    vp = ctypes.c_void_p
    lib.rwkv_init_from_file.argtypes = [ctypes.c_char_p, ctypes.c_uint32, ctypes.c_uint32]
    lib.rwkv_init_from_file.restype = vp
    lib.rwkv_clone_context.argtypes = [vp, ctypes.c_uint32]
    lib.rwkv_clone_context.restype = vp
    lib.rwkv_eval.argtypes = [vp, ctypes.c_uint32, P_FLOAT, P_FLOAT, P_FLOAT]
    lib.rwkv_eval.restype = ctypes.c_bool
    lib.rwkv_eval_sequence.argtypes = [vp, P_U32, ctypes.c_size_t, P_FLOAT, P_FLOAT, P_FLOAT]
    lib.rwkv_eval_sequence.restype = ctypes.c_bool
    lib.rwkv_eval_sequence_in_chunks.argtypes = [vp, P_U32, ctypes.c_size_t, ctypes.c_size_t, P_FLOAT, P_FLOAT, P_FLOAT]
    lib.rwkv_eval_sequence_in_chunks.restype = ctypes.c_bool
    for name in ("rwkv_get_n_vocab", "rwkv_get_n_embed", "rwkv_get_n_layer", "rwkv_get_state_len", "rwkv_get_logits_len"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = ctypes.c_size_t
    for name in ("rwkv_get_state_buffer_element_count", "rwkv_get_logits_buffer_element_count"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = ctypes.c_uint32
    lib.rwkv_init_state.argtypes = [vp, P_FLOAT]
    lib.rwkv_init_state.restype = None
    lib.rwkv_free.argtypes = [vp]
    lib.rwkv_free.restype = None
    lib.rwkv_quantize_model_file.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
    lib.rwkv_quantize_model_file.restype = ctypes.c_bool
    lib.rwkv_get_system_info_string.argtypes = []
    lib.rwkv_get_system_info_string.restype = ctypes.c_char_p
    lib.rwkv_set_print_errors.argtypes = [vp, ctypes.c_bool]
    lib.rwkv_set_print_errors.restype = None
    lib.rwkv_get_print_errors.argtypes = [vp]
    lib.rwkv_get_print_errors.restype = ctypes.c_bool
    lib.rwkv_get_last_error.argtypes = [vp]
    lib.rwkv_get_last_error.restype = ctypes.c_int
    return lib


def load_reference_library():
    path = reference_library_path()
    if path is None:
        raise FileNotFoundError("oracle/_ref/librwkv_ref*.so not built; run `make -C oracle` where /root/reference exists")
    return bind_rwkv_api(ctypes.CDLL(path))
