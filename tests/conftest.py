import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")
VERSIONS = ["4v0-660K", "5v1-730K", "5v2-730K", "6v0-3m", "7v0-834K"]
FILE_FORMATS = ["FP32", "FP16", "Q5_0", "Q5_1"]
QUANT_FORMATS = ["Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q8_0"]
PROMPT = list(b'"in')                                   # tests/logit_difference_validator.inc:48-49
LONG_PROMPT = list(b"This is a port of [BlinkDL/RWKV-LM](https://github.com/BlinkDL/RWKV-LM")   # tests/test_eval_sequence_in_chunks.c:69

# Known-answer tables of the reference's own tests: |sum(logits - expected)| must stay within 1.05x of these.
# tests/test_tiny_rwkv.c:38-54
DIFF_SUM_FULL = {"4v0-660K": (0.001, -0.013652), "5v1-730K": (0.001, -0.289921), "5v2-730K": (0.001, 0.455912),
                 "6v0-3m": (0.001, -0.416620), "7v0-834K": (0.001, 0.005766)}
# tests/test_tiny_rwkv.c:70-101 (FP32 -> Q) and :103-134 (FP16 -> Q); order Q4_0, Q4_1, Q5_0, Q5_1, Q8_0
DIFF_SUM_Q_FROM_FP32 = {
    "4v0-660K": (-0.160030, -0.547409, -0.170404, 0.278034, 0.076282),
    "5v1-730K": (117.932594, -26.712271, -163.439407, -18.017435, 0.585238),
    "5v2-730K": (35.271305, 67.015076, 25.273308, 48.068733, -9.441034),
    "6v0-3m": (-7.588121, 21.939022, -27.332073, 3.576909, -9.539596),
    "7v0-834K": (0.136785, 0.002614, -0.063645, -0.064663, 0.011924),
}
DIFF_SUM_Q_FROM_FP16 = {
    "4v0-660K": (0.154614, -0.539827, -0.180142, 0.294953, 0.077226),
    "5v1-730K": (119.471931, -28.245888, -159.870956, -39.708530, -0.962695),
    "5v2-730K": (34.135971, 65.573822, 21.588751, 29.726818, -7.242277),
    "6v0-3m": (-7.660988, 21.797060, -27.269241, 3.405264, -9.734720),
    "7v0-834K": (0.136678, -0.005140, -0.064447, -0.063531, 0.010921),
}
# tests/test_quantization_format_compatibility.c:22-35 (checked-in Q5_0 / Q5_1 files; no 7v0 there)
DIFF_SUM_CHECKED_IN = {"4v0-660K": (-0.170404, 0.278034), "5v1-730K": (-163.439407, -18.017435),
                       "5v2-730K": (25.273308, 48.068733), "6v0-3m": (-21.151785, 3.576909)}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def model_path(ver, fmt):
    return os.path.join(GOLDEN, "models", f"tiny-rwkv-{ver}-{fmt}.bin")


def expected_logits(ver):
    return np.fromfile(os.path.join(GOLDEN, "logits", f"expected-logits-{ver}.bin"), dtype=np.float32)


@pytest.fixture(scope="session")
def pkg():
    import __graft_entry__
    return __graft_entry__.load_package()


@pytest.fixture(scope="session")
def lib(pkg):
    """The product library. Tests never fall back to anything else if it is missing."""
    if not os.path.isfile(pkg.library_path()):
        import __graft_entry__
        __graft_entry__.build()
    lib = pkg.load_rwkv_shared_library()
    lib.rwkv_set_print_errors(None, False)
    return lib


@pytest.fixture(scope="session")
def ref_outputs():
    return np.load(os.path.join(GOLDEN, "ref_outputs.npz"))


@pytest.fixture(scope="session")
def quantized_dir(lib, tmp_path_factory):
    """FP32->Q and FP16->Q variants of every fixture, made by OUR quantizer (as tests/test_tiny_rwkv.c:136-171 does)."""
    d = tmp_path_factory.mktemp("quantized")
    for ver in VERSIONS:
        for src in ("FP32", "FP16"):
            for fmt in QUANT_FORMATS:
                lib.rwkv_quantize_model_file(model_path(ver, src), str(d / f"tiny-rwkv-{ver}-{src}-to-{fmt}.bin"), fmt)
    return d


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
