"""Generates tests/golden/converter_sha256.json: SHA-256 of the files the REFERENCE converter (python/convert_pytorch_to_ggml.py,
imported from /root/reference, so run this where the reference checkout exists) writes for the seeded synthetic checkpoints of
tools/synthetic_checkpoint.py. tests/test_converter.py checks that our converter produces the same bytes."""
import hashlib
import importlib.util
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import synthetic_checkpoint as sc  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_convert", "/root/reference/python/convert_pytorch_to_ggml.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

out = {}
with tempfile.TemporaryDirectory() as d:
    for kind in sc.SHAPES:
        for dtype in ("FP32", "FP16"):
            path = os.path.join(d, f"{kind}-{dtype}.bin")
            ref.write_state_dict(sc.make_state_dict(kind, seed=7), path, dtype)
            data = open(path, "rb").read()
            out[f"{kind}/{dtype}"] = {"sha256": hashlib.sha256(data).hexdigest(), "bytes": len(data)}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "converter_sha256.json"), "w"), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
