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


def test_reference_known_answer_bytes(converter, tmp_path):
    """The known-answer vector of the reference's own converter test (python/convert_pytorch_to_ggml.test.py:11-48): a 3 x 2 embedding
    and one 1-element LayerNorm weight must serialise to exactly these bytes."""
    import struct
    import torch
    state_dict = {"emb.weight": torch.tensor([[1, 2], [3, 4], [5, 6]], dtype=torch.float32),
                  "blocks.0.ln1.weight": torch.tensor([1], dtype=torch.float32)}
    path = str(tmp_path / "known.bin")
    converter.write_state_dict(state_dict, dest_path=path, data_type="FP32")
    expected = struct.pack("=iiiiii" + "iiiii10sffffff" + "iiii19sf",
                           0x67676D66, 101, 3, 2, 1, 0,
                           2, 10, 0, 2, 3, b"emb.weight", 1.0, 2.0, 3.0, 4.0, 5.0, 6.0,
                           1, 19, 0, 1, b"blocks.0.ln1.weight", 1.0)
    assert open(path, "rb").read() == expected
