"""Golden outputs of the compiled reference (oracle/_ref) for the large-activation robustness test
(tests/test_gpu_parity.py::test_large_activations_*): 40 tokens from a token-shift state of 1e3 on the Q5_0 fixtures -- the largest
power of ten at which the reference itself still produces finite logits (from 1e4 on, and from 1e3 on for Q5_1 / FP16 files, its
fp16 conversions overflow and every logit is NaN; measured by this script, printed below).

    python tests/golden/make_large_act_ref.py        # needs oracle/_ref (make -C oracle)
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_lib  # noqa: E402

PF = ctypes.POINTER(ctypes.c_float)
PU = ctypes.POINTER(ctypes.c_uint32)
TOKENS = [(7919 * i + 3) % 256 for i in range(40)]
CASES = (("6v0-3m", 128, 12), ("5v2-730K", 64, 12))


def big_state(n, C, L, mag):
    state = np.zeros(n, np.float32)
    per = n // L
    for layer in range(L):
        state[layer * per: layer * per + 2 * C] = mag
    return state


def run(lib, path, C, L, mag, how):
    ctx = lib.rwkv_init_from_file(path.encode(), 1, 0)
    n, V = lib.rwkv_get_state_len(ctx), lib.rwkv_get_logits_len(ctx)
    state = big_state(n, C, L, mag)
    logits, out = np.zeros(V, np.float32), np.zeros(n, np.float32)
    if how == "sequence":
        ta = np.array(TOKENS, np.uint32)
        assert lib.rwkv_eval_sequence(ctx, ta.ctypes.data_as(PU), len(TOKENS), state.ctypes.data_as(PF), out.ctypes.data_as(PF), logits.ctypes.data_as(PF))
    else:
        cur = state
        for t in TOKENS:
            out = np.zeros(n, np.float32)
            assert lib.rwkv_eval(ctx, t, cur.ctypes.data_as(PF), out.ctypes.data_as(PF), logits.ctypes.data_as(PF))
            cur = out
    lib.rwkv_free(ctx)
    return logits, out


def main():
    libs = {}
    for name in ("librwkv_ref.so", "librwkv_ref_native.so"):
        p = os.path.join(ref_lib.REF_DIR, name)
        if os.path.isfile(p):
            lib = ctypes.CDLL(p)
            ref_lib.bind_rwkv_api(lib)
            libs[name] = lib
    golden = {}
    for ver, C, L in CASES:
        for fmt in ("Q5_0", "Q5_1", "FP16"):
            path = os.path.join(ROOT, "tests", "golden", "models", f"tiny-rwkv-{ver}-{fmt}.bin")
            for mag in (1e3, 1e4, 3e5):
                res = {k: run(lib, path, C, L, mag, "sequence")[0] for k, lib in libs.items()}
                fin = {k: bool(np.isfinite(v).all()) for k, v in res.items()}
                print(ver, fmt, mag, "finite:", fin)
        path = os.path.join(ROOT, "tests", "golden", "models", f"tiny-rwkv-{ver}-Q5_0.bin")
        outs = []
        for k, lib in libs.items():
            for how in ("sequence", "serial"):
                outs.append(run(lib, path, C, L, 1e3, how))
        spread_l = max(float(np.abs(a[0] - outs[0][0]).max()) for a in outs)
        spread_s = max(float(np.abs(a[1] - outs[0][1]).max()) for a in outs)
        print(ver, "Q5_0 1e3: reference self-spread (builds x serial/sequence): logits", spread_l, "state", spread_s, "max|logit|", float(np.abs(outs[0][0]).max()))
        golden[ver + "/logits"], golden[ver + "/state"] = outs[0]
        golden[ver + "/spread"] = np.array([spread_l, spread_s], np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "large_act_ref.npz"), **golden)


if __name__ == "__main__":
    main()
