"""Regenerates tests/golden/ref_outputs.npz from the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile
from /root/reference). Run in the build container: python tests/golden/make_ref_outputs.py

For every fixture model and every on-the-fly quantised variant the reference's own test builds
(tests/test_tiny_rwkv.c:136-171) it stores the logits after the prompt `"in` (serial rwkv_eval x3, fresh
state); for the 20 checked-in files also the final state and the outputs after a 70-token prompt.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_lib  # noqa: E402

VERSIONS = ["4v0-660K", "5v1-730K", "5v2-730K", "6v0-3m", "7v0-834K"]
PROMPT = list(b'"in')
LONG_PROMPT = list(b"This is a port of [BlinkDL/RWKV-LM](https://github.com/BlinkDL/RWKV-LM")  # tests/test_eval_sequence_in_chunks.c:69


def main():
    ref = ref_lib.load_reference_library()
    ref.rwkv_set_print_errors(None, False)
    PF = ref_lib.P_FLOAT

    def run(path, toks):
        ctx = ref.rwkv_init_from_file(path.encode(), 2, 0)
        assert ctx, path
        n = ref.rwkv_get_state_len(ctx)
        st, lg = np.zeros(n, np.float32), np.zeros(256, np.float32)
        ref.rwkv_init_state(ctx, st.ctypes.data_as(PF))
        for t in toks:
            assert ref.rwkv_eval(ctx, t, st.ctypes.data_as(PF), st.ctypes.data_as(PF), lg.ctypes.data_as(PF))
        ref.rwkv_free(ctx)
        return lg, st

    out = {}
    tmp = tempfile.mkdtemp()
    for ver in VERSIONS:
        for fmt in ["FP32", "FP16", "Q5_0", "Q5_1"]:
            path = f"{ROOT}/tests/golden/models/tiny-rwkv-{ver}-{fmt}.bin"
            lg, st = run(path, PROMPT)
            out[f"{ver}/{fmt}/logits"], out[f"{ver}/{fmt}/state"] = lg, st
            lg, st = run(path, LONG_PROMPT)
            out[f"{ver}/{fmt}/long_logits"], out[f"{ver}/{fmt}/long_state"] = lg, st
        for src in ["FP32", "FP16"]:
            for fmt in ["Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q8_0"]:
                q = f"{tmp}/{ver}-{src}-to-{fmt}.bin"
                assert ref.rwkv_quantize_model_file(f"{ROOT}/tests/golden/models/tiny-rwkv-{ver}-{src}.bin".encode(), q.encode(), fmt.encode())
                lg, _ = run(q, PROMPT)
                out[f"{ver}/{src}-to-{fmt}/logits"] = lg
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ref_outputs.npz"), **out)
    print("wrote", len(out), "arrays; library:", ref_lib.reference_library_path())


if __name__ == "__main__":
    main()
