"""rwkv_eval / rwkv_eval_sequence with the host state copied per layer group on copy streams (rwkv_b200_set_overlap) must give
exactly the bytes of the plain upload - evaluate - download order, for every architecture, with aliased state buffers, NULL
state_in / state_out / logits_out, pageable and pinned memory. Also: CUDA tensors as
state / logits buffers of the Python wrapper."""
import ctypes

import numpy as np
import pytest

from conftest import LONG_PROMPT, VERSIONS, model_path

pytestmark = pytest.mark.gpu
P_F = ctypes.POINTER(ctypes.c_float)


def run(lib, ctx, toks, overlap, n_state, n_logits, skip_logits=False, seq=0):
    lib.library.rwkv_b200_set_overlap(ctx.ptr, overlap)
    state = np.zeros(n_state, dtype=np.float32)
    logits = np.zeros(n_logits, dtype=np.float32)
    i = 0
    while i < len(toks):
        want = (not skip_logits) or i + max(seq, 1) >= len(toks)
        sin = None if i == 0 else state.ctypes.data_as(P_F)          # aliased in / out from the second call on
        if seq:
            chunk = toks[i:i + seq]
            arr = (ctypes.c_uint32 * len(chunk))(*chunk)
            assert lib.library.rwkv_eval_sequence(ctx.ptr, arr, len(chunk), sin, state.ctypes.data_as(P_F), logits.ctypes.data_as(P_F) if want else None)
            i += len(chunk)
        else:
            assert lib.library.rwkv_eval(ctx.ptr, toks[i], sin, state.ctypes.data_as(P_F), logits.ctypes.data_as(P_F) if want else None)
            i += 1
    return logits.copy(), state.copy()


@pytest.mark.parametrize("ver", VERSIONS)
@pytest.mark.parametrize("fmt", ["FP32", "Q5_1"])
def test_overlapped_copies_bitwise(lib, ver, fmt):
    ctx = lib.rwkv_init_from_file(model_path(ver, fmt), 1, 0)
    try:
        n_state, n_logits = lib.rwkv_get_state_buffer_element_count(ctx), lib.rwkv_get_logits_buffer_element_count(ctx)
        toks = LONG_PROMPT[:20]
        want = run(lib, ctx, toks, False, n_state, n_logits)
        for skip in (False, True):
            got = run(lib, ctx, toks, True, n_state, n_logits, skip_logits=skip)
            assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes(), (ver, fmt, skip)
        want_seq = run(lib, ctx, toks, False, n_state, n_logits, seq=7)
        got_seq = run(lib, ctx, toks, True, n_state, n_logits, seq=7)
        assert got_seq[0].tobytes() == want_seq[0].tobytes() and got_seq[1].tobytes() == want_seq[1].tobytes()
        assert want_seq[1].tobytes() == want[1].tobytes()          # and sequence mode == serial, as ever
        # state_out == NULL (logits only) and state_in == NULL with overlap on
        lib.library.rwkv_b200_set_overlap(ctx.ptr, True)
        lg = np.zeros(n_logits, dtype=np.float32)
        assert lib.library.rwkv_eval(ctx.ptr, toks[0], None, None, lg.ctypes.data_as(P_F))
        st = np.zeros(n_state, dtype=np.float32)
        assert lib.library.rwkv_eval(ctx.ptr, toks[0], None, st.ctypes.data_as(P_F), None)
        ref_l, ref_s = run(lib, ctx, toks[:1], False, n_state, n_logits)
        assert lg.tobytes() == ref_l.tobytes() and st.tobytes() == ref_s.tobytes()
    finally:
        lib.rwkv_free(ctx)


@pytest.mark.parametrize("ver", VERSIONS)
def test_pageable_buffers_through_the_bounce_path_bitwise(lib, ver):
    """Pageable caller memory (numpy arrays) of any size through the pinned bounce buffers + helper threads (engine.cu:
    eval_host_overlapped; normally from 4 MiB of state on): the same bytes as the plain order, with aliased state_in == state_out,
    NULL state_in, logits skipped, serial and sequence calls."""
    ctx = lib.rwkv_init_from_file(model_path(ver, "Q5_1"), 1, 0)
    try:
        n_state, n_logits = lib.rwkv_get_state_buffer_element_count(ctx), lib.rwkv_get_logits_buffer_element_count(ctx)
        toks = LONG_PROMPT[:24]
        want = run(lib, ctx, toks, False, n_state, n_logits)
        want_seq = run(lib, ctx, toks, False, n_state, n_logits, seq=9)
        lib.library.rwkv_b200_set_bounce_min_bytes(0)
        try:
            for skip in (False, True):
                got = run(lib, ctx, toks, True, n_state, n_logits, skip_logits=skip)
                assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes(), (ver, skip)
            got_seq = run(lib, ctx, toks, True, n_state, n_logits, seq=9)
            assert got_seq[0].tobytes() == want_seq[0].tobytes() and got_seq[1].tobytes() == want_seq[1].tobytes()
            # distinct in / out buffers, and a state handed back without logits
            s_in, s_out = want[1].copy(), np.zeros(n_state, dtype=np.float32)
            assert lib.library.rwkv_eval(ctx.ptr, toks[0], s_in.ctypes.data_as(P_F), s_out.ctypes.data_as(P_F), None)
            lib.library.rwkv_b200_set_overlap(ctx.ptr, False)
            s_ref = np.zeros(n_state, dtype=np.float32)
            assert lib.library.rwkv_eval(ctx.ptr, toks[0], s_in.ctypes.data_as(P_F), s_ref.ctypes.data_as(P_F), None)
            assert s_out.tobytes() == s_ref.tobytes() and s_in.tobytes() == want[1].tobytes()
        finally:
            lib.library.rwkv_b200_set_bounce_min_bytes(4 << 20)
    finally:
        lib.rwkv_free(ctx)


def test_overlap_with_pinned_buffers_real_head_size(pkg, lib, tmp_path):
    import torch
    import synthetic_model as sm
    path = str(tmp_path / "mid.bin")
    sm.write_direct(path, "rwkv6-mid", "Q5_1", seed=2)
    ctx = lib.rwkv_init_from_file(path, 1, 0)
    try:
        n_state, n_logits = lib.rwkv_get_state_buffer_element_count(ctx), lib.rwkv_get_logits_buffer_element_count(ctx)
        toks = sm.synthetic_tokens(6, n_logits)
        want = run(lib, ctx, toks, False, n_state, n_logits)
        state = torch.zeros(n_state, dtype=torch.float32).pin_memory()
        logits = torch.zeros(n_logits, dtype=torch.float32).pin_memory()
        lib.library.rwkv_b200_set_overlap(ctx.ptr, True)
        for i, t in enumerate(toks):
            sp = ctypes.cast(state.data_ptr(), P_F)
            assert lib.library.rwkv_eval(ctx.ptr, t, None if i == 0 else sp, sp, ctypes.cast(logits.data_ptr(), P_F))
        assert logits.numpy().tobytes() == want[0].tobytes() and state.numpy().tobytes() == want[1].tobytes()
    finally:
        lib.rwkv_free(ctx)


def test_cuda_tensors_as_state_and_logits(pkg, lib):
    import torch
    m = pkg.RWKVModel(lib, model_path("6v0-3m", "Q5_1"), thread_count=1)
    try:
        toks = LONG_PROMPT[:10]
        st, want_logits = None, None
        for t in toks:
            want_logits, st = m.eval(t, st, use_numpy=True)
        dstate, dlogits = None, None
        for t in toks:
            dlogits, dstate = m.eval(t, dstate) if dstate is not None else m.eval(t, None, torch.zeros(m.state_len, device="cuda"), torch.zeros(m.n_vocab, device="cuda"))
        assert dstate.device.type == "cuda" and dlogits.device.type == "cuda"
        assert dlogits.cpu().numpy().tobytes() == want_logits.tobytes() and dstate.cpu().numpy().tobytes() == st.tobytes()
    finally:
        m.free()
