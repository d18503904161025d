"""rwkv.cpp_b200/convert_pytorch_to_ggml.py against the reference converter (python/convert_pytorch_to_ggml.py): the same seeded
synthetic PyTorch checkpoints must give the same bytes (golden hashes made by tests/golden/make_converter_golden.py, which imports
the reference converter), and the converted files must load and evaluate (numpy oracle here; the GPU run is in
tests/test_gpu_converter.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN

import rwkv_oracle as ro
import synthetic_checkpoint as sc

GOLD = json.load(open(os.path.join(GOLDEN, "converter_sha256.json")))


@pytest.fixture(scope="module")
def converter(pkg):
    import importlib
    return importlib.import_module("rwkv_cpp_b200.convert_pytorch_to_ggml")


@pytest.mark.parametrize("kind", list(sc.SHAPES))
@pytest.mark.parametrize("dtype", ["FP32", "FP16"])
def test_same_bytes_as_reference_converter(converter, tmp_path, kind, dtype):
    path = str(tmp_path / f"{kind}-{dtype}.bin")
    converter.write_state_dict(sc.make_state_dict(kind, seed=7), path, dtype)
    data = open(path, "rb").read()
    assert len(data) == GOLD[f"{kind}/{dtype}"]["bytes"]
    assert hashlib.sha256(data).hexdigest() == GOLD[f"{kind}/{dtype}"]["sha256"]


@pytest.mark.parametrize("kind", list(sc.SHAPES))
def test_converted_file_evaluates_in_oracle(converter, tmp_path, kind):
    path = str(tmp_path / f"{kind}.bin")
    converter.write_state_dict(sc.make_state_dict(kind, seed=3), path, "FP32")
    m = ro.OracleModel(path)
    p = sc.SHAPES[kind]
    assert (m.n_vocab, m.n_embed, m.n_layer) == (p["V"], p["C"], p["L"])
    logits, state = m.eval_sequence([1, 2, 3, 4], None)
    assert np.isfinite(logits).all() and np.isfinite(state).all() and logits.std() > 0
    # the float16 name aliases of the reference CLI
    converter.write_state_dict(sc.make_state_dict(kind, seed=3), path, "float16")
    assert ro.OracleModel(path).n_layer == p["L"]


def test_architecture_detection_and_errors(converter, tmp_path):
    assert converter.detect_architecture(sc.make_state_dict("v4")) == (4, 0)
    assert converter.detect_architecture(sc.make_state_dict("v5.1")) == (5, 1)
    assert converter.detect_architecture(sc.make_state_dict("v5.2")) == (5, 2)
    assert converter.detect_architecture(sc.make_state_dict("v6")) == (6, 0)
    assert converter.detect_architecture(sc.make_state_dict("v7")) == (7, 0)
    with pytest.raises(ValueError):
        converter.write_state_dict(sc.make_state_dict("v4"), str(tmp_path / "x.bin"), "Q5_1")
    with pytest.raises(ValueError):
        converter.count_layers({"emb.weight": None})
