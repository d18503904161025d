"""Pipeline-stage entry points (include/rwkv_b200.h) on ONE GPU: a model cut into two stage contexts, chained through a
device buffer, must give exactly the logits of the whole model -- token by token (serial path, CUDA graphs) and for one sequence pass."""
import ctypes

import numpy as np
import pytest

from conftest import LONG_PROMPT, VERSIONS, model_path

pytestmark = pytest.mark.gpu
PF = ctypes.POINTER(ctypes.c_float)
PU = ctypes.POINTER(ctypes.c_uint32)


@pytest.mark.parametrize("ver", VERSIONS)
@pytest.mark.parametrize("fmt", ["FP32", "Q5_1"])
def test_two_stages_on_one_gpu_match_the_whole_model(pkg, lib, ver, fmt):
    import torch
    L = lib.library
    path = model_path(ver, fmt)
    whole = pkg.RWKVModel(lib, path, thread_count=1)
    n_layer, n_vocab = lib.rwkv_get_n_layer(whole._ctx), lib.rwkv_get_logits_len(whole._ctx)
    cut = pkg.pipeline.stage_layers(n_layer, 2, 0)[1]
    a = lib.rwkv_b200_init_from_file_ex(path, 0, 0, cut)
    b = lib.rwkv_b200_init_from_file_ex(path, 0, cut, -1)
    try:
        for c in (a, b):
            assert L.rwkv_b200_state_load(c.ptr, None)
        tokens = LONG_PROMPT[:12]
        n_hidden = L.rwkv_b200_stage_hidden_len(a.ptr, 1)
        assert n_hidden == pkg.pipeline.hidden_floats(lib.rwkv_get_n_embed(whole._ctx), 1, 7 if ver.startswith("7") else 6)
        buf = torch.zeros(n_hidden, dtype=torch.float32, device="cuda:0")
        got = np.zeros(n_vocab, np.float32)
        state = None
        for t in tokens:       # serial: the third token on replays the captured graphs of both stages
            want, state = whole.eval(t, state, use_numpy=True)
            tok = (ctypes.c_uint32 * 1)(t)
            assert L.rwkv_b200_stage_eval(a.ptr, tok, 1, None, ctypes.c_void_p(buf.data_ptr()), False, None)
            assert L.rwkv_b200_synchronize(a.ptr)
            assert L.rwkv_b200_stage_eval(b.ptr, None, 1, ctypes.c_void_p(buf.data_ptr()), None, True, None)
            assert L.rwkv_b200_stage_logits(b.ptr, got.ctypes.data_as(PF), None)
            assert np.array_equal(got, want), (ver, fmt, t)
        # one sequence pass of 9 tokens continuing from the same states
        seq = LONG_PROMPT[12:21]
        want, state = whole.eval_sequence(seq, state, use_numpy=True)
        arr = (ctypes.c_uint32 * len(seq))(*seq)
        buf = torch.zeros(L.rwkv_b200_stage_hidden_len(a.ptr, len(seq)), dtype=torch.float32, device="cuda:0")
        assert L.rwkv_b200_stage_eval(a.ptr, arr, len(seq), None, ctypes.c_void_p(buf.data_ptr()), False, None)
        assert L.rwkv_b200_synchronize(a.ptr)
        assert L.rwkv_b200_stage_eval(b.ptr, None, len(seq), ctypes.c_void_p(buf.data_ptr()), None, True, None)
        assert L.rwkv_b200_stage_logits(b.ptr, got.ctypes.data_as(PF), None)
        assert np.array_equal(got, want)
        # argument checks: a later stage needs activations, an earlier one an output buffer
        assert not L.rwkv_b200_stage_eval(b.ptr, arr, 1, None, None, True, None)
        assert not L.rwkv_b200_stage_eval(a.ptr, arr, 1, None, None, False, None)
    finally:
        lib.rwkv_free(a); lib.rwkv_free(b); whole.free()
