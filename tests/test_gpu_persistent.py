"""The persistent single-token kernel (csrc/kernels/decode_persistent.cu) against the per-launch path it replaces: the two must
agree BIT FOR BIT (logits and state) because the reference's tests memcmp serial against sequence / chunked evaluation
(tests/test_eval_sequence_in_chunks.c:54) and only the serial path runs the persistent kernel."""
import numpy as np
import pytest

from conftest import LONG_PROMPT, PROMPT, model_path

pytestmark = pytest.mark.gpu

PERSISTENT_VERSIONS = ["5v1-730K", "5v2-730K", "6v0-3m"]


def run_serial(lib, m, tokens, persistent):
    lib.library.rwkv_b200_set_persistent(m._ctx.ptr, persistent)
    state, all_logits = None, []
    for t in tokens:
        logits, state = m.eval(t, state, use_numpy=True)
        all_logits.append(logits.copy())
    return np.stack(all_logits), state.copy()


@pytest.mark.parametrize("ver", PERSISTENT_VERSIONS)
@pytest.mark.parametrize("fmt", ["FP32", "FP16", "Q5_0", "Q5_1"])
def test_persistent_equals_per_launch_bitwise_fixtures(pkg, lib, ver, fmt):
    m = pkg.RWKVModel(lib, model_path(ver, fmt), thread_count=1)
    try:
        toks = LONG_PROMPT[:24]
        want_logits, want_state = run_serial(lib, m, toks, False)
        got_logits, got_state = run_serial(lib, m, toks, True)
        assert lib.library.rwkv_b200_persistent_state(m._ctx.ptr) == 1, "the fixture shape must fit the persistent kernel"
        assert np.isfinite(got_logits).all()
        assert got_logits.tobytes() == want_logits.tobytes(), np.abs(got_logits - want_logits).max()
        assert got_state.tobytes() == want_state.tobytes(), np.abs(got_state - want_state).max()
        # and the reference's invariant: serial (persistent kernel) == sequence mode (per-launch kernels)
        seq_logits, seq_state = m.eval_sequence(toks, None, use_numpy=True)
        assert seq_logits.tobytes() == got_logits[-1].tobytes() and seq_state.tobytes() == got_state.tobytes()
    finally:
        m.free()


@pytest.mark.parametrize("ver", PERSISTENT_VERSIONS)
def test_persistent_quantized_on_the_fly(pkg, lib, quantized_dir, ver):
    for fmt in ("Q4_0", "Q4_1", "Q8_0"):
        m = pkg.RWKVModel(lib, str(quantized_dir / f"tiny-rwkv-{ver}-FP32-to-{fmt}.bin"), thread_count=1)
        try:
            want = run_serial(lib, m, PROMPT * 3, False)
            got = run_serial(lib, m, PROMPT * 3, True)
            assert lib.library.rwkv_b200_persistent_state(m._ctx.ptr) == 1
            assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes(), (ver, fmt)
        finally:
            m.free()


@pytest.mark.parametrize("preset,fmt", [("rwkv6-small", "Q5_1"), ("rwkv6-small", "Q8_0"), ("rwkv6-small", "FP16"), ("rwkv5-small", "Q4_0"),
                                        ("rwkv5.1-small", "Q5_0"), ("rwkv6-mid", "Q5_1"), ("rwkv6-wide", "Q5_1"), ("rwkv6-wide", "Q4_0")])
def test_persistent_equals_per_launch_bitwise_real_head_size(pkg, lib, tmp_path, preset, fmt):
    """Head size 64, LoRA ranks and FFN widths of real checkpoints; `rwkv6-mid` has rows long enough (ffn 7168) to be split over
    several warps and n_embed 2048 (two channels per LayerNorm thread); `rwkv6-wide` is one layer of the 7B shape."""
    import synthetic_model as sm
    path = str(tmp_path / f"{preset}-{fmt}.bin")
    sm.write_direct(path, preset, fmt, seed=5)
    m = pkg.RWKVModel(lib, path, thread_count=1)
    try:
        toks = sm.synthetic_tokens(12, m.n_vocab)
        want = run_serial(lib, m, toks, False)
        got = run_serial(lib, m, toks, True)
        assert lib.library.rwkv_b200_persistent_state(m._ctx.ptr) == 1
        assert np.isfinite(got[0]).all()
        assert got[0].tobytes() == want[0].tobytes(), np.abs(got[0] - want[0]).max()
        assert got[1].tobytes() == want[1].tobytes(), np.abs(got[1] - want[1]).max()
    finally:
        m.free()


def test_persistent_logits_skipping_and_clones(pkg, lib):
    """Both programs of a context (with / without the head) and two contexts of one model evaluated alternately."""
    import ctypes
    path = model_path("6v0-3m", "Q5_1")
    a = lib.rwkv_init_from_file(path, 1, 0)
    b = lib.rwkv_clone_context(a, 1)
    n_state, n_logits = lib.rwkv_get_state_buffer_element_count(a), lib.rwkv_get_logits_buffer_element_count(a)
    P_F = ctypes.POINTER(ctypes.c_float)

    def run(ctx, toks, persistent, skip):
        lib.library.rwkv_b200_set_persistent(ctx.ptr, persistent)
        state = np.zeros(n_state, dtype=np.float32)
        logits = np.zeros(n_logits, dtype=np.float32)
        first = True
        for i, t in enumerate(toks):
            want = (not skip) or i == len(toks) - 1
            ok = lib.library.rwkv_eval(ctx.ptr, t, None if first else state.ctypes.data_as(P_F), state.ctypes.data_as(P_F),
                                       logits.ctypes.data_as(P_F) if want else None)
            assert ok
            first = False
        return logits.copy(), state.copy()

    toks = LONG_PROMPT[:16]
    want = run(a, toks, False, False)
    for ctx in (a, b):
        for skip in (False, True):
            got = run(ctx, toks, True, skip)
            assert got[0].tobytes() == want[0].tobytes() and got[1].tobytes() == want[1].tobytes()
    # interleaved: one token on a, one on b
    lib.library.rwkv_b200_set_persistent(a.ptr, True)
    lib.library.rwkv_b200_set_persistent(b.ptr, True)
    sa = np.zeros(n_state, dtype=np.float32); sb = np.zeros(n_state, dtype=np.float32)
    la = np.zeros(n_logits, dtype=np.float32); lb = np.zeros(n_logits, dtype=np.float32)
    for i, t in enumerate(toks):
        assert lib.library.rwkv_eval(a.ptr, t, None if i == 0 else sa.ctypes.data_as(P_F), sa.ctypes.data_as(P_F), la.ctypes.data_as(P_F))
        assert lib.library.rwkv_eval(b.ptr, t, None if i == 0 else sb.ctypes.data_as(P_F), sb.ctypes.data_as(P_F), lb.ctypes.data_as(P_F))
    assert la.tobytes() == want[0].tobytes() and lb.tobytes() == want[0].tobytes() and sa.tobytes() == want[1].tobytes() and sb.tobytes() == want[1].tobytes()
    lib.rwkv_free(b)
    lib.rwkv_free(a)
