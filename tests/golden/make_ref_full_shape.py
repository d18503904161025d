"""Regenerates tests/golden/ref_full_shape.npz: outputs of the UNMODIFIED reference (oracle/_ref, built by oracle/Makefile
from /root/reference) on BASELINE.json's configurations at their FULL shapes. Run in the build container (needs ~20 GB of
RAM and a few minutes of CPU):  python tests/golden/make_ref_full_shape.py [case ...]

The model files are synthetic (tools/synthetic_model.py: seeded, so the GPU box regenerates the identical bytes); every
case stores the tokens it ran, the logits of the last token (float32) and a strided sample of the final state, for
  decode   n single-token rwkv_eval calls from a fresh state (state_in aliases state_out, like the reference's tests)
  chunk    one rwkv_eval_sequence_in_chunks call (chunk = 128) over the same kind of token stream
The GPU tests (tests/test_gpu_full_shape.py) compare the CUDA path with these numbers at the same tolerances as the tiny
fixtures. Only logits + a state sample are kept so the fixture stays ~2 MB.
"""
import ctypes
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ref_lib  # noqa: E402
import synthetic_model as sm  # noqa: E402

STATE_STRIDE = 1009          # prime: the sample walks through every slot kind of every layer
# name: (preset, format, seed, decode tokens, chunk tokens)
CASES = {
    "rwkv4-169m:Q5_1": ("rwkv4-169m", "Q5_1", 3, 64, 64),        # BASELINE.json config 1
    "rwkv5-1b5:Q4_0": ("rwkv5-1b5", "Q4_0", 3, 8, 128),          # config 2 (chunk = 128 prefill)
    "rwkv7-2b9:FP16": ("rwkv7-2b9", "FP16", 3, 8, 0),            # config 3
    "rwkv6-7b:Q8_0": ("rwkv6-7b", "Q8_0", 1, 4, 0),              # config 4's file (pipelined across GPUs in the bench)
    "rwkv6-7b:Q5_1": ("rwkv6-7b", "Q5_1", 1, 4, 128),            # the north-star workload (bench.py's file, seed 1)
}


def main():
    names = sys.argv[1:] or list(CASES)
    out_path = os.path.join(ROOT, "tests", "golden", "ref_full_shape.npz")
    out = dict(np.load(out_path)) if os.path.isfile(out_path) else {}
    ref = ref_lib.load_reference_library()
    ref.rwkv_set_print_errors(None, False)
    PF, PU = ref_lib.P_FLOAT, ref_lib.P_U32
    threads = int(os.environ.get("RWKV_REF_THREADS", "8"))
    tmp = os.environ.get("RWKV_B200_BENCH_DIR", tempfile.gettempdir())
    for name in names:
        preset, fmt, seed, n_decode, n_chunk = CASES[name]
        path = os.path.join(tmp, f"{preset}-{fmt}-seed{seed}.bin")
        if not os.path.isfile(path):
            sm.write_direct(path, preset, fmt, seed=seed)
        t0 = time.time()
        ctx = ref.rwkv_init_from_file(path.encode(), threads, 0)
        assert ctx, path
        n_state, n_vocab = ref.rwkv_get_state_len(ctx), ref.rwkv_get_logits_len(ctx)
        toks = sm.synthetic_tokens(max(n_decode, n_chunk), n_vocab)
        st, lg = np.zeros(n_state, np.float32), np.zeros(n_vocab, np.float32)
        ref.rwkv_init_state(ctx, st.ctypes.data_as(PF))
        for t in toks[:n_decode]:
            assert ref.rwkv_eval(ctx, t, st.ctypes.data_as(PF), st.ctypes.data_as(PF), lg.ctypes.data_as(PF))
        out[f"{name}/decode_tokens"] = np.asarray(toks[:n_decode], np.uint32)
        out[f"{name}/decode_logits"] = lg.copy()
        out[f"{name}/decode_state_sample"] = st[::STATE_STRIDE].copy()
        if n_chunk:
            arr = (ctypes.c_uint32 * n_chunk)(*toks[:n_chunk])
            st2, lg2 = np.zeros(n_state, np.float32), np.zeros(n_vocab, np.float32)
            assert ref.rwkv_eval_sequence_in_chunks(ctx, arr, n_chunk, 128, None, st2.ctypes.data_as(PF), lg2.ctypes.data_as(PF))
            out[f"{name}/chunk_tokens"] = np.asarray(toks[:n_chunk], np.uint32)
            out[f"{name}/chunk_logits"] = lg2.copy()
            out[f"{name}/chunk_state_sample"] = st2[::STATE_STRIDE].copy()
        ref.rwkv_free(ctx)
        print(f"{name}: {n_decode} decode + {n_chunk} chunk tokens in {time.time() - t0:.1f}s, |logits|max {np.abs(lg).max():.3f}", flush=True)
        np.savez_compressed(out_path, **out)
    print("wrote", out_path, os.path.getsize(out_path), "bytes; library:", ref_lib.reference_library_path())


if __name__ == "__main__":
    main()
