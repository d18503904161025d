"""How far does the UNMODIFIED reference differ from ITSELF on the full-shape synthetic files? Runs the two builds of oracle/_ref
(x86-64-v3 = AVX2 kernels, and -march=native = AVX-512 where the host has it; same sources, different vector widths, i.e. different fp32
summation orders) on the cases of tests/golden/make_ref_full_shape.py and prints max|logits_a - logits_b| per case, for the decode
run and for the chunk run, plus the same after ONE token. The bars of tests/test_gpu_full_shape.py are set from these numbers:
random-init weights amplify the rounding flips of the Q8 / fp16 activation quantisation far more than trained checkpoints do.
    python tools/ref_self_spread.py [case ...]        (build container only: needs oracle/_ref and ~20 GB of RAM)"""
import ctypes, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tools")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import ref_lib
import synthetic_model as sm
from make_ref_full_shape import CASES

def run(libpath, path, n_decode, n_chunk, threads):
    lib = ctypes.CDLL(libpath); ref_lib.bind_rwkv_api(lib)
    lib.rwkv_set_print_errors(None, False)
    PF = ref_lib.P_FLOAT
    ctx = lib.rwkv_init_from_file(path.encode(), threads, 0); assert ctx
    n_state, n_vocab = lib.rwkv_get_state_len(ctx), lib.rwkv_get_logits_len(ctx)
    toks = sm.synthetic_tokens(max(n_decode, n_chunk), n_vocab)
    st, lg = np.zeros(n_state, np.float32), np.zeros(n_vocab, np.float32)
    lib.rwkv_init_state(ctx, st.ctypes.data_as(PF))
    per_tok = []
    for t in toks[:n_decode]:
        assert lib.rwkv_eval(ctx, t, st.ctypes.data_as(PF), st.ctypes.data_as(PF), lg.ctypes.data_as(PF))
        per_tok.append(lg.copy())
    out = {"decode": per_tok, "state": st.copy()}
    if n_chunk:
        arr = (ctypes.c_uint32 * n_chunk)(*toks[:n_chunk])
        st2, lg2 = np.zeros(n_state, np.float32), np.zeros(n_vocab, np.float32)
        assert lib.rwkv_eval_sequence_in_chunks(ctx, arr, n_chunk, 128, None, st2.ctypes.data_as(PF), lg2.ctypes.data_as(PF))
        out["chunk"] = lg2
    lib.rwkv_free(ctx)
    return out

def main():
    names = sys.argv[1:] or list(CASES)
    a_path, b_path = os.path.join(ref_lib.REF_DIR, "librwkv_ref.so"), os.path.join(ref_lib.REF_DIR, "librwkv_ref_native.so")
    tmp = os.environ.get("RWKV_B200_BENCH_DIR", "/tmp/rwkv_b200_bench"); os.makedirs(tmp, exist_ok=True)
    threads = int(os.environ.get("RWKV_REF_THREADS", "8"))
    res = {}
    for name in names:
        preset, fmt, seed, n_decode, n_chunk = CASES[name]
        path = os.path.join(tmp, f"{preset}-{fmt}-seed{seed}.bin")
        if not os.path.isfile(path): sm.write_direct(path, preset, fmt, seed=seed)
        a, b = run(a_path, path, n_decode, n_chunk, threads), run(b_path, path, n_decode, n_chunk, threads)
        r = {"decode_tokens": n_decode, "logits_absmax": float(np.abs(a["decode"][-1]).max()),
             "decode_spread_per_token": [float(np.abs(x - y).max()) for x, y in zip(a["decode"], b["decode"])],
             "state_spread": float(np.abs(a["state"] - b["state"]).max())}
        if n_chunk: r["chunk_spread"] = float(np.abs(a["chunk"] - b["chunk"]).max())
        res[name] = r
        print(name, json.dumps(r), flush=True)
    json.dump(res, open(os.path.join(ROOT, "tests", "golden", "ref_self_spread.json"), "w"), indent=1)
main()
