"""Interactive GPU parity sweep (not part of the test suite): every fixture x format, ours vs the compiled
reference (oracle/_ref) and vs the golden logits; serial, sequence and chunked modes."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import __graft_entry__ as g
import ref_lib
pkg = g.load_package(); lib = pkg.load_rwkv_shared_library()
ref = ref_lib.load_reference_library(); ref.rwkv_set_print_errors(None, False)
PF = ref_lib.P_FLOAT
print(lib.rwkv_get_system_info_string())
prompt = list(b'"in'); long_prompt = [(7919 * i + 13) % 256 for i in range(70)]

def ref_run(path, toks):
    ctx = ref.rwkv_init_from_file(path.encode(), 2, 0)
    n = ref.rwkv_get_state_len(ctx); st = np.zeros(n, np.float32); lg = np.zeros(256, np.float32)
    ref.rwkv_init_state(ctx, st.ctypes.data_as(PF))
    for t in toks:
        ref.rwkv_eval(ctx, t, st.ctypes.data_as(PF), st.ctypes.data_as(PF), lg.ctypes.data_as(PF))
    ref.rwkv_free(ctx); return lg, st

import tempfile
tmp = tempfile.mkdtemp()
bad = 0
for ver in ["4v0-660K", "5v1-730K", "5v2-730K", "6v0-3m", "7v0-834K"]:
    exp = np.fromfile(f"{ROOT}/tests/golden/logits/expected-logits-{ver}.bin", dtype=np.float32)
    files = [(fmt, f"{ROOT}/tests/golden/models/tiny-rwkv-{ver}-{fmt}.bin") for fmt in ["FP32", "FP16", "Q5_0", "Q5_1"]]
    for fmt in ["Q4_0", "Q4_1", "Q8_0"]:
        out = f"{tmp}/{ver}-{fmt}.bin"
        lib.rwkv_set_print_errors(None, False)
        lib.rwkv_quantize_model_file(f"{ROOT}/tests/golden/models/tiny-rwkv-{ver}-FP32.bin", out, fmt)
        files.append(("FP32>" + fmt, out))
    for fmt, path in files:
        try:
            m = pkg.RWKVModel(lib, path, thread_count=1)
        except Exception as e:
            print(ver, fmt, "LOAD FAILED", e); bad += 1; continue
        s = None
        for t in prompt:
            lg, s = m.eval(t, s, use_numpy=True)
        lg = lg.copy(); s = s.copy()
        lq, sq = m.eval_sequence(prompt, None, use_numpy=True)
        rl, rs = ref_run(path, prompt)
        # long prompt: serial vs chunks
        s2 = None
        for t in long_prompt:
            l2, s2 = m.eval(t, s2, use_numpy=True)
        l2 = l2.copy(); s2 = s2.copy()
        same = []
        for ch in (1, 2, 8, 10, 70):
            l3, s3 = m.eval_sequence_in_chunks(long_prompt, None, chunk_size=ch, use_numpy=True)
            same.append(bool(np.array_equal(l3, l2) and np.array_equal(s3, s2)))
        rl2, rs2 = ref_run(path, long_prompt)
        m.free()
        fin = np.isfinite(lg).all()
        print(f"{ver:9s} {fmt:10s} vs_ref logits {np.abs(lg-rl).max():.2e} state {np.abs(s-rs).max():.2e} | seq==serial {np.array_equal(lg,lq) and np.array_equal(s,sq)} | "
              f"vs_golden {np.abs(lg-exp).max():.2e} diffsum {float((lg-exp).sum()):+.4f} (ref {float((rl-exp).sum()):+.4f}) | long(70) vs_ref {np.abs(l2-rl2).max():.2e} st {np.abs(s2-rs2).max():.2e} chunks_bitexact {same} finite {fin}")
print("done")
