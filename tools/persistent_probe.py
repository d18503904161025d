"""Diagnoses the persistent single-token kernel against the per-launch path on one model file.

    python tools/persistent_probe.py <model.bin> [--tokens 6] [--trace]

Prints, per layer, whether the slots of the recurrent state written by each phase agree bit for bit (the state holds
LN1(x) = att_xx, the WKV state and LN2(x) = ffn_xx of every layer, so the first differing slot names the first phase that
went wrong), the logits difference, timings of both paths and (--trace) the persistent kernel's phase timeline.
"""
import argparse
import ctypes
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--tokens", type=int, default=6)
    ap.add_argument("--trace", action="store_true")
    ap.add_argument("--time", type=int, default=0, help="time this many resident decode steps on both paths")
    ap.add_argument("--batch", type=int, default=0, help="time this many steps of batched multi-sequence decode for 2, 4, 8, 16 sequences")
    ap.add_argument("--e2e", type=int, default=0, help="time this many rwkv_eval calls with pinned host state for persistent x overlap")
    args = ap.parse_args()
    pkg = __graft_entry__.load_package()
    lib = pkg.load_rwkv_shared_library()
    L = lib.library
    ctx = lib.rwkv_init_from_file(args.model, 1, 0)
    n_state, n_logits = lib.rwkv_get_state_buffer_element_count(ctx), lib.rwkv_get_logits_buffer_element_count(ctx)
    n_layer, C = L.rwkv_get_n_layer(ctx.ptr), L.rwkv_get_n_embed(ctx.ptr)
    per_layer = n_state // n_layer
    P_F = ctypes.POINTER(ctypes.c_float)
    toks = [(7919 * i + 3) % n_logits for i in range(args.tokens)]

    def run(persistent):
        L.rwkv_b200_set_persistent(ctx.ptr, persistent)
        state = np.zeros(n_state, dtype=np.float32)
        logits = np.zeros(n_logits, dtype=np.float32)
        states = []
        for i, t in enumerate(toks):
            ok = L.rwkv_eval(ctx.ptr, t, None if i == 0 else state.ctypes.data_as(P_F), state.ctypes.data_as(P_F), logits.ctypes.data_as(P_F))
            if not ok:
                print(f"rwkv_eval failed at token {i} (persistent={persistent}), flags 0x{lib.rwkv_get_last_error(ctx):x}")
                sys.exit(2)
            states.append(state.copy())
        return logits.copy(), states

    want_logits, want_states = run(False)
    got_logits, got_states = run(True)
    print("persistent_state:", L.rwkv_b200_persistent_state(ctx.ptr))
    bad = False
    for i, (ws, gs) in enumerate(zip(want_states, got_states)):
        if ws.tobytes() == gs.tobytes():
            continue
        bad = True
        print(f"token {i}: state differs")
        for l in range(n_layer):
            w, g = ws[l * per_layer:(l + 1) * per_layer], gs[l * per_layer:(l + 1) * per_layer]
            slots = (("ffn_xx", 0, C), ("att_xx", C, 2 * C), ("wkv", 2 * C, per_layer))
            msg = []
            for name, a, b in slots:
                d = np.abs(w[a:b] - g[a:b])
                nbad = int((w[a:b].view(np.uint32) != g[a:b].view(np.uint32)).sum())
                if nbad:
                    msg.append(f"{name}: {nbad}/{b - a} differ, max {d.max():.3e}, first at {int(np.argmax(w[a:b].view(np.uint32) != g[a:b].view(np.uint32)))}")
            if msg:
                print(f"  layer {l}: " + "; ".join(msg))
        break
    dl = np.abs(want_logits - got_logits)
    print(f"logits: max |diff| {dl.max():.3e}, bitwise equal {want_logits.tobytes() == got_logits.tobytes()}, finite {bool(np.isfinite(got_logits).all())}")
    print("RESULT", "MISMATCH" if bad or want_logits.tobytes() != got_logits.tobytes() else "BIT-EXACT")

    if args.time:
        arr = (ctypes.c_uint32 * (args.time + 8))(*[(7919 * i) % n_logits for i in range(args.time + 8)])
        for persistent in (False, True):
            L.rwkv_b200_set_persistent(ctx.ptr, persistent)
            L.rwkv_b200_state_load(ctx.ptr, None)
            ms = L.rwkv_b200_time_resident(ctx.ptr, arr, 1, args.time, 8, True)
            print(f"resident decode, persistent={persistent}: {ms / args.time:.4f} ms/token")
    if args.e2e:
        import torch
        st = torch.zeros(n_state, dtype=torch.float32).pin_memory()
        lg = torch.zeros(n_logits, dtype=torch.float32).pin_memory()
        sp, lp = ctypes.cast(st.data_ptr(), P_F), ctypes.cast(lg.data_ptr(), P_F)
        for persistent in (False, True):
            for overlap in (False, True):
                L.rwkv_b200_set_persistent(ctx.ptr, persistent)
                L.rwkv_b200_set_overlap(ctx.ptr, overlap)
                lib.rwkv_init_state(ctx, st.data_ptr())
                for i in range(6):
                    L.rwkv_eval(ctx.ptr, (7919 * i) % n_logits, sp, sp, lp)
                t0 = time.perf_counter()
                for i in range(args.e2e):
                    L.rwkv_eval(ctx.ptr, (7919 * (i + 6)) % n_logits, sp, sp, lp)
                dt = (time.perf_counter() - t0) / args.e2e * 1e3
                print(f"e2e rwkv_eval, pinned host state, persistent={persistent} overlap={overlap}: {dt:.4f} ms/token ({1e3 / dt:.1f} tok/s)  checksum {float(lg.sum()):.6f}")
        L.rwkv_b200_set_overlap(ctx.ptr, False)
    if args.batch:
        for n_seq in (2, 4, 8, 16):
            bptr = L.rwkv_b200_batch_create(ctx.ptr, n_seq)
            if not bptr:
                print(f"batch of {n_seq}: not supported for this model")
                break
            b = ctypes.c_void_p(bptr)
            toks_b = (ctypes.c_uint32 * n_seq)(*[(7919 * i + 11) % n_logits for i in range(n_seq)])
            for _ in range(3):
                L.rwkv_b200_batch_eval(b, toks_b, True)
            L.rwkv_b200_synchronize(b)
            t0 = time.perf_counter()
            for _ in range(args.batch):
                L.rwkv_b200_batch_eval(b, toks_b, True)
            L.rwkv_b200_synchronize(b)
            dt = (time.perf_counter() - t0) / args.batch * 1e3
            print(f"batched decode, {n_seq} sequences: {dt:.3f} ms/step, {n_seq / dt * 1e3:.0f} tok/s aggregate")
            L.rwkv_free(b)
    if args.trace:
        L.rwkv_b200_set_persistent(ctx.ptr, True)
        buf = (ctypes.c_double * 4096)()
        L.rwkv_b200_phase_trace(ctx.ptr, buf, 4096)      # arm
        one = (ctypes.c_uint32 * 1)(toks[0])
        for _ in range(3):
            L.rwkv_b200_eval_resident(ctx.ptr, one, 1, True, None)
        n = L.rwkv_b200_phase_trace(ctx.ptr, buf, 4096)
        b = [buf[i] for i in range(max(n, 0))]
        print(f"phase trace: {n} boundaries, total {b[-1] if b else -1:.1f} us")
        if n > 8:
            dur = np.diff(np.array(b))
            per = 7 if (n - 2) % 7 == 0 else 5
            body = dur[:(len(dur) - 1) // per * per].reshape(-1, per)
            print("per-layer phase durations (us), median over layers:", np.round(np.median(body, axis=0), 2).tolist(), "layer total", round(float(np.median(body.sum(axis=1))), 2))
            print("first layer:", np.round(body[0], 2).tolist(), " head:", round(float(dur[-1]), 2))
            mk = (ctypes.c_double * (3 * 700))()
            nm = L.rwkv_b200_phase_marks(ctx.ptr, mk, 700)
            if nm > 0:
                marks = np.array([mk[i] for i in range(3 * nm)]).reshape(nm, 3)
                starts = np.array(b[:nm])
                rel = np.where(marks >= 0, marks - starts[:, None], np.nan)       # since the phase's barrier released
                nl = (nm - 1) // per
                relb = rel[:nl * per].reshape(nl, per, 3)
                with np.errstate(all="ignore"):
                    med = np.nanmedian(relb, axis=0)
                print("CTA 0, us after the phase's barrier released [inputs ready, column staged, tiles consumed], median over layers:")
                for i in range(per):
                    print(f"  phase {i}: {np.round(med[i], 2).tolist()}  (phase length {np.round(np.median(body[:, i]), 2)})")
    lib.rwkv_free(ctx)


if __name__ == "__main__":
    main()
