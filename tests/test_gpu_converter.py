"""Converter -> quantizer -> engine, end to end on the GPU: a synthetic PyTorch checkpoint converted by
rwkv.cpp_b200/convert_pytorch_to_ggml.py (optionally quantised by rwkv_quantize_model_file) evaluates like the numpy oracle
on the same file."""
import importlib

import numpy as np
import pytest

import rwkv_oracle as ro
import synthetic_checkpoint as sc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,dtype,quant,tol", [("v4", "FP32", None, 2e-4), ("v5.1", "FP32", None, 2e-4), ("v5.2", "FP16", None, 2e-2),
                                                  ("v6", "FP32", None, 2e-4), ("v7", "FP32", None, 2e-4), ("v6", "FP16", "Q5_1", 2e-1),
                                                  ("v7", "FP16", "Q8_0", 2e-1)])
def test_converted_checkpoint_matches_oracle(pkg, lib, tmp_path, kind, dtype, quant, tol):
    conv = importlib.import_module("rwkv_cpp_b200.convert_pytorch_to_ggml")
    path = str(tmp_path / f"{kind}-{dtype}.bin")
    conv.write_state_dict(sc.make_state_dict(kind, seed=9), path, dtype)
    if quant:
        qpath = str(tmp_path / f"{kind}-{quant}.bin")
        lib.rwkv_quantize_model_file(path, qpath, quant)
        path = qpath
    toks = [1, 7, 3, 49, 0, 12]
    want_logits, want_state = ro.OracleModel(path).eval_sequence(toks, None)
    m = pkg.RWKVModel(lib, path, thread_count=1)
    try:
        logits, state = m.eval_sequence(toks, None, use_numpy=True)
    finally:
        m.free()
    assert np.isfinite(logits).all()
    assert np.abs(logits - want_logits).max() <= tol, np.abs(logits - want_logits).max()
    assert np.abs(state - want_state).max() <= 50 * tol
