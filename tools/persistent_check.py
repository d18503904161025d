"""Bit-exactness check of the persistent single-token kernel against the per-launch path, for one or more model files, in a
process of its own (tests/test_zz_gpu_persistent.py runs it as a subprocess: a fault inside a cooperative kernel poisons the CUDA
context of the process it happens in, which must not be the pytest process).

    python tools/persistent_check.py [--tokens N] [--overlap] [--skip-logits] [--clones] model.bin [model.bin ...]

Prints one line per model: `RESULT <path> BIT-EXACT|MISMATCH|NOT-PERSISTENT|EVAL-FAILED|INIT-FAILED` and exits 0 only if every
model was evaluated by the persistent kernel and agreed bit for bit with the per-launch kernels.
"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402

P_F = ctypes.POINTER(ctypes.c_float)


def run(L, ctx, toks, n_state, n_logits, persistent, overlap, skip_logits):
    L.rwkv_b200_set_persistent(ctx, persistent)
    L.rwkv_b200_set_overlap(ctx, overlap)
    state = np.zeros(n_state, dtype=np.float32)
    logits = np.zeros(n_logits, dtype=np.float32)
    all_logits = []
    for i, t in enumerate(toks):
        want = (not skip_logits) or i == len(toks) - 1
        ok = L.rwkv_eval(ctx, t, None if i == 0 else state.ctypes.data_as(P_F), state.ctypes.data_as(P_F), logits.ctypes.data_as(P_F) if want else None)
        if not ok:
            return None
        if want:
            all_logits.append(logits.copy())
    return np.stack(all_logits), state.copy()


def check(lib, path, args):
    L = lib.library
    try:
        ctx_obj = lib.rwkv_init_from_file(path, 1, 0)
    except ValueError:
        return "INIT-FAILED"
    ctx = ctx_obj.ptr
    try:
        n_state, n_logits = lib.rwkv_get_state_buffer_element_count(ctx_obj), lib.rwkv_get_logits_buffer_element_count(ctx_obj)
        toks = [(7919 * i + 3) % n_logits for i in range(args.tokens)]
        want = run(L, ctx, toks, n_state, n_logits, False, False, False)
        if want is None:
            return "EVAL-FAILED"
        variants = [(False, False)]
        if args.overlap:
            variants.append((True, False))
        if args.skip_logits:
            variants.append((args.overlap, True))
        for overlap, skip in variants:
            got = run(L, ctx, toks, n_state, n_logits, True, overlap, skip)
            if got is None:
                return "EVAL-FAILED"
            if L.rwkv_b200_persistent_state(ctx) != 1:
                return "NOT-PERSISTENT"
            ref_logits = want[0][-1:] if skip else want[0]
            if got[0].tobytes() != ref_logits.tobytes() or got[1].tobytes() != want[1].tobytes():
                return "MISMATCH"
        if args.clones:     # a second context of the same model, evaluated alternately with the first
            clone_obj = lib.rwkv_clone_context(ctx_obj, 1)
            clone = clone_obj.ptr
            L.rwkv_b200_set_persistent(ctx, True)
            L.rwkv_b200_set_persistent(clone, True)
            sa, sb = np.zeros(n_state, np.float32), np.zeros(n_state, np.float32)
            la, lb = np.zeros(n_logits, np.float32), np.zeros(n_logits, np.float32)
            for i, t in enumerate(toks):
                oka = L.rwkv_eval(ctx, t, None if i == 0 else sa.ctypes.data_as(P_F), sa.ctypes.data_as(P_F), la.ctypes.data_as(P_F))
                okb = L.rwkv_eval(clone, t, None if i == 0 else sb.ctypes.data_as(P_F), sb.ctypes.data_as(P_F), lb.ctypes.data_as(P_F))
                if not (oka and okb):
                    return "EVAL-FAILED"
            lib.rwkv_free(clone_obj)
            if la.tobytes() != want[0][-1].tobytes() or lb.tobytes() != want[0][-1].tobytes() or sa.tobytes() != want[1].tobytes() or sb.tobytes() != want[1].tobytes():
                return "MISMATCH"
        # the reference's invariant on top: serial (persistent kernel) == sequence mode (per-launch kernels)
        L.rwkv_b200_set_persistent(ctx, False)
        arr = (ctypes.c_uint32 * len(toks))(*toks)
        s2, l2 = np.zeros(n_state, np.float32), np.zeros(n_logits, np.float32)
        if not L.rwkv_eval_sequence(ctx, arr, len(toks), None, s2.ctypes.data_as(P_F), l2.ctypes.data_as(P_F)):
            return "EVAL-FAILED"
        if s2.tobytes() != want[1].tobytes() or l2.tobytes() != want[0][-1].tobytes():
            return "MISMATCH"
        return "BIT-EXACT"
    finally:
        lib.rwkv_free(ctx_obj)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("models", nargs="+")
    ap.add_argument("--tokens", type=int, default=12)
    ap.add_argument("--overlap", action="store_true", help="also with the host state copied per layer group (rwkv_b200_set_overlap)")
    ap.add_argument("--skip-logits", action="store_true", help="also with logits requested for the last token only")
    ap.add_argument("--clones", action="store_true", help="also two contexts of the model evaluated alternately")
    args = ap.parse_args()
    pkg = __graft_entry__.load_package()
    lib = pkg.load_rwkv_shared_library()
    lib.rwkv_set_print_errors(None, True)
    bad = 0
    for path in args.models:
        verdict = check(lib, path, args)
        print(f"RESULT {path} {verdict}", flush=True)
        bad += verdict != "BIT-EXACT"
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
