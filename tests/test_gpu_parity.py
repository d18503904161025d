"""GPU parity tests proper: the CUDA path, called through the C ABI, against the compiled reference's outputs
(tests/golden/ref_outputs.npz), the reference's golden logits and known-answer tables, the numpy oracle on larger
synthetic shapes, and the reference's bit-exactness invariants (its tests memcmp states)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import (DIFF_SUM_CHECKED_IN, DIFF_SUM_FULL, DIFF_SUM_Q_FROM_FP16, DIFF_SUM_Q_FROM_FP32, FILE_FORMATS, LONG_PROMPT, PROMPT,
                      QUANT_FORMATS, ROOT, VERSIONS, expected_logits, model_path)

pytestmark = pytest.mark.gpu

# max-abs tolerance vs the compiled reference on the same file. FP32 is the north star's 1e-3 bar (we sit at ~1e-6).
# FP16 / quantised: the reference differs from ITSELF by 1.5e-3 (6v0 FP16) and 8.3e-3 (5v1 Q5_0) between its AVX2 and
# AVX-512 builds because fp16 / int8 activation rounding flips amplify fp32 summation-order noise (parity);
# the bars below are ~3x those measured spreads.
TOL = {"FP32": 1e-3, "FP16": 5e-3, "Q": 5e-2}
# The 70-token prompt: rounding flips of the Q8 / fp16 activation quantisation accumulate in the recurrent state. Measured in the build
# container: the numpy oracle (pinned to the reference at 2e-5 on the 3-token prompt; same arithmetic, different fp32 summation
# order) differs from the compiled reference after 70 tokens by up to 3.3e-3 (FP16: 5v1) and 5.3e-2 / 0.20 in logits / state
# (Q5_1: 5v1 / 6v0) -- so that is how far two faithful implementations sit apart; the bars are ~3x those spreads.
LONG_TOL = {"FP32": 1e-3, "FP16": 2e-2, "Q": 1.5e-1}
HELLO = list(b"hello world")


def serial(model, tokens, want_logits=True):
    state, logits = None, None
    for t in tokens:
        logits, state = model.eval(t, state, use_numpy=True)
    return logits.copy(), state.copy()


@pytest.fixture(scope="module")
def models(pkg, lib):
    cache = {}

    def get(path):
        if path not in cache:
            cache[path] = pkg.RWKVModel(lib, str(path), thread_count=2)
        return cache[path]
    yield get
    for m in cache.values():
        m.free()


@pytest.mark.parametrize("ver", VERSIONS)
@pytest.mark.parametrize("fmt", FILE_FORMATS)
def test_fixture_logits_and_state(models, ref_outputs, ver, fmt):
    m = models(model_path(ver, fmt))
    logits, state = serial(m, PROMPT)
    tol = TOL.get(fmt, TOL["Q"])
    assert np.isfinite(logits).all()
    assert np.abs(logits - ref_outputs[f"{ver}/{fmt}/logits"]).max() <= tol
    assert np.abs(state - ref_outputs[f"{ver}/{fmt}/state"]).max() <= 10 * tol
    if fmt == "FP32":   # the golden vector itself
        assert np.abs(logits - expected_logits(ver)).max() <= 1e-3
    seq_logits, seq_state = m.eval_sequence(PROMPT, None, use_numpy=True)
    assert np.array_equal(seq_logits, logits) and np.array_equal(seq_state, state)   # serial == sequence, bit for bit


@pytest.mark.parametrize("ver", VERSIONS)
def test_reference_diff_sum_tables(models, quantized_dir, ver):
    """tests/test_tiny_rwkv.c + tests/test_quantization_format_compatibility.c: |sum(logits - expected)| <= 1.05 |table|,
    serial and sequence mode, for FP32, FP16, the checked-in Q5 files and all ten on-the-fly quantised variants."""
    exp = expected_logits(ver)
    cases = [(model_path(ver, "FP32"), DIFF_SUM_FULL[ver][0]), (model_path(ver, "FP16"), DIFF_SUM_FULL[ver][1])]
    if ver in DIFF_SUM_CHECKED_IN:
        cases += [(model_path(ver, "Q5_0"), DIFF_SUM_CHECKED_IN[ver][0]), (model_path(ver, "Q5_1"), DIFF_SUM_CHECKED_IN[ver][1])]
    for i, fmt in enumerate(QUANT_FORMATS):
        cases.append((quantized_dir / f"tiny-rwkv-{ver}-FP32-to-{fmt}.bin", DIFF_SUM_Q_FROM_FP32[ver][i]))
        cases.append((quantized_dir / f"tiny-rwkv-{ver}-FP16-to-{fmt}.bin", DIFF_SUM_Q_FROM_FP16[ver][i]))
    for path, budget in cases:
        m = models(path)
        logits, _ = serial(m, PROMPT)
        assert abs(float((logits - exp).sum())) <= abs(budget) * 1.05, (str(path), float((logits - exp).sum()), budget)
        logits, _ = m.eval_sequence(PROMPT, None, use_numpy=True)
        assert abs(float((logits - exp).sum())) <= abs(budget) * 1.05, (str(path), "sequence")


@pytest.mark.parametrize("ver", VERSIONS)
def test_quantized_on_the_fly_vs_reference(models, quantized_dir, ref_outputs, ver):
    for src in ("FP32", "FP16"):
        for fmt in QUANT_FORMATS:
            m = models(quantized_dir / f"tiny-rwkv-{ver}-{src}-to-{fmt}.bin")
            logits, _ = serial(m, PROMPT)
            assert np.abs(logits - ref_outputs[f"{ver}/{src}-to-{fmt}/logits"]).max() <= TOL["Q"], (ver, src, fmt)


@pytest.mark.parametrize("ver", VERSIONS)
@pytest.mark.parametrize("fmt", ["FP32", "Q5_1"])
def test_chunked_equals_serial_bitwise(models, ref_outputs, ver, fmt):
    """tests/test_eval_sequence_in_chunks.c:45-54 (the reference runs it on 5v2 FP32 only; here every arch, two formats)."""
    m = models(model_path(ver, fmt))
    for prompt in (LONG_PROMPT, LONG_PROMPT[:1]):
        want_logits, want_state = serial(m, prompt)
        for chunk in (1, 2, 8, 10, 31, 64, 1000):
            logits, state = m.eval_sequence_in_chunks(prompt, None, chunk_size=chunk, use_numpy=True)
            if fmt == "FP32" or min(chunk, len(prompt)) < 32:
                # batch-invariant SIMT path: bit for bit
                assert state.tobytes() == want_state.tobytes(), (ver, fmt, chunk)
                assert logits.tobytes() == want_logits.tobytes(), (ver, fmt, chunk)
            else:
                # passes of >= 32 tokens of non-F32 weights run on the tensor cores (fp16 operands holding the reference's Q8
                # activation values): not bit-identical to the dp4a path, but within the bar we hold against the reference
                assert np.abs(logits - want_logits).max() <= LONG_TOL["Q"], (ver, fmt, chunk, np.abs(logits - want_logits).max())
    if fmt == "FP32":   # after 70 tokens FP32 still tracks the reference closely
        logits, state = serial(m, LONG_PROMPT)
        assert np.abs(logits - ref_outputs[f"{ver}/FP32/long_logits"]).max() <= 1e-3
        assert np.abs(state - ref_outputs[f"{ver}/FP32/long_state"]).max() <= 1e-2


@pytest.mark.parametrize("ver", VERSIONS)
@pytest.mark.parametrize("fmt", ["FP16", "Q5_0", "Q5_1"])
def test_long_prompt_vs_reference_all_paths(models, ref_outputs, ver, fmt):
    """The 70-token prompt of tests/test_eval_sequence_in_chunks.c against the compiled reference's outputs on the same file
    (tests/golden/ref_outputs.npz: */long_logits, */long_state): serial, one sequence call (tensor-core path for non-F32
    weights) and chunks of 32 / 64 -- the reference computes all of them identically (memcmp, :54), so each must sit within
    the same bar (LONG_TOL: what separates two faithful implementations after 70 tokens). This pins the >= 32-token path to the
    reference."""
    m = models(model_path(ver, fmt))
    tol = LONG_TOL.get(fmt, LONG_TOL["Q"])
    want_l, want_s = ref_outputs[f"{ver}/{fmt}/long_logits"], ref_outputs[f"{ver}/{fmt}/long_state"]
    runs = {"serial": serial(m, LONG_PROMPT), "sequence": m.eval_sequence(LONG_PROMPT, None, use_numpy=True)}
    for chunk in (32, 64):
        runs[f"chunks of {chunk}"] = m.eval_sequence_in_chunks(LONG_PROMPT, None, chunk_size=chunk, use_numpy=True)
    for how, (logits, state) in runs.items():
        assert np.isfinite(logits).all() and np.isfinite(state).all(), (ver, fmt, how)
        el, es = np.abs(logits - want_l).max(), np.abs(state - want_s).max()
        assert el <= tol, (ver, fmt, how, el)
        assert es <= 5 * tol, (ver, fmt, how, es)


def _big_shift_state(m, C, L, mag):
    state = np.zeros(m.state_len, np.float32)
    per_layer = state.size // L
    for layer in range(L):
        state[layer * per_layer: layer * per_layer + 2 * C] = mag
    return state


def test_large_activations_stay_finite_on_the_tensor_core_path(pkg, lib):
    """activations beyond the fp16 range (65504) must not turn into inf / NaN on the >= 32-token path of quantised
    weights: the fp16 operands of the tensor-core GEMM carry a per-token power-of-two scale that the epilogue takes out again.
    The carried token-shift state is set to 3e5, so the mixed inputs of the first token of every matrix exceed 65504.
    Only finiteness is asserted here: the compiled reference itself returns NaN logits in this regime (from a shift state of 1e4 on
    for Q5_0 files, of 1e3 on for Q5_1 / FP16 files: tests/golden/make_large_act_ref.py prints the table), so there is no reference
    value to agree with; the magnitude at which the reference is still finite is compared in the next test."""
    toks = [(7919 * i + 3) % 256 for i in range(40)]
    for ver, C, L in (("6v0-3m", 128, 12), ("5v2-730K", 64, 12)):
        m = pkg.RWKVModel(lib, model_path(ver, "Q5_0"), thread_count=1)
        try:
            state = _big_shift_state(m, C, L, 3.0e5)
            l_seq, s_seq = m.eval_sequence(toks, state.copy(), use_numpy=True)
            st, lg = state.copy(), None
            for t in toks:
                lg, st = m.eval(t, st, use_numpy=True)
            assert np.isfinite(l_seq).all() and np.isfinite(s_seq).all(), ver
            assert np.isfinite(lg).all(), ver
        finally:
            m.free()


def test_large_activations_match_the_reference_where_it_is_finite(pkg, lib):
    """Shift state 1e3 on the Q5_0 fixtures: the largest power of ten the compiled reference survives. Both of our paths -- 40 serial
    rwkv_eval calls (dp4a GEMV) and one 40-token rwkv_eval_sequence (tcgen05 GEMM, scaled fp16 operands) -- against the reference's
    outputs (tests/golden/large_act_ref.npz, written by make_large_act_ref.py). Bar: the long-prompt bar of the quantised formats
    (LONG_TOL, 1.5e-1 on logits of magnitude 6 - 16): in this scenario the numpy oracle -- the reference's arithmetic with another fp32
    summation order -- sits 1.7e-2 (6v0) / 3.7e-2 (5v2) away from the compiled reference after the 40 tokens, 7e2 of 1.3e6 in the 6v0
    state (measured in the build container), which is how far two faithful implementations are apart here."""
    ref = np.load(os.path.join(str(ROOT), "tests", "golden", "large_act_ref.npz"))
    toks = [(7919 * i + 3) % 256 for i in range(40)]
    for ver, C, L in (("6v0-3m", 128, 12), ("5v2-730K", 64, 12)):
        want_l, want_s, spread = ref[ver + "/logits"], ref[ver + "/state"], ref[ver + "/spread"]
        tol_l = max(LONG_TOL["Q"], 10 * float(spread[0]))
        tol_s = max(5 * tol_l, 10 * float(spread[1]), 2e-3 * float(np.abs(want_s).max()))
        m = pkg.RWKVModel(lib, model_path(ver, "Q5_0"), thread_count=1)
        try:
            state = _big_shift_state(m, C, L, 1.0e3)
            l_seq, s_seq = m.eval_sequence(toks, state.copy(), use_numpy=True)
            st, lg = state.copy(), None
            for t in toks:
                lg, st = m.eval(t, st, use_numpy=True)
            for how, l, s_ in (("sequence", l_seq, s_seq), ("serial", lg, st)):
                assert np.isfinite(l).all() and np.isfinite(s_).all(), (ver, how)
                assert np.abs(l - want_l).max() <= tol_l, (ver, how, float(np.abs(l - want_l).max()), tol_l)
                assert np.abs(s_ - want_s).max() <= tol_s, (ver, how, float(np.abs(s_ - want_s).max()), tol_s)
        finally:
            m.free()


@pytest.mark.parametrize("ver", VERSIONS)
def test_logits_skipping_keeps_state(lib, pkg, ver):
    """tests/test_logit_calculation_skipping.c: logits_out == NULL must not change the state (serial and sequence)."""
    ctx = lib.rwkv_init_from_file(model_path(ver, "FP32"), 2, 0)
    n, v = lib.rwkv_get_state_len(ctx), lib.rwkv_get_logits_len(ctx)
    want, got, logits = np.zeros(n, np.float32), np.zeros(n, np.float32), np.zeros(v, np.float32)
    lib.rwkv_eval(ctx, HELLO[0], None, want.ctypes.data, logits.ctypes.data)
    for t in HELLO[1:]:
        lib.rwkv_eval(ctx, t, want.ctypes.data, want.ctypes.data, logits.ctypes.data)   # state_in aliases state_out
    lib.rwkv_eval(ctx, HELLO[0], None, got.ctypes.data, None)
    for t in HELLO[1:]:
        lib.rwkv_eval(ctx, t, got.ctypes.data, got.ctypes.data, None)
    assert got.tobytes() == want.tobytes()
    got[:] = 0
    lib.rwkv_eval_sequence(ctx, HELLO, None, got.ctypes.data, None)
    assert got.tobytes() == want.tobytes()
    lib.rwkv_free(ctx)


def test_clone_outlives_parent(lib):
    """tests/test_context_cloning.c:10-57."""
    path = model_path("5v2-730K", "FP32")
    ctx = lib.rwkv_init_from_file(path, 2, 0)
    n, v = lib.rwkv_get_state_len(ctx), lib.rwkv_get_logits_len(ctx)
    state, want = np.zeros(n, np.float32), np.zeros(v, np.float32)
    lib.rwkv_eval(ctx, HELLO[0], None, state.ctypes.data, want.ctypes.data)
    for t in HELLO[1:]:
        lib.rwkv_eval(ctx, t, state.ctypes.data, state.ctypes.data, want.ctypes.data)
    clone = lib.rwkv_clone_context(ctx, 2)
    lib.rwkv_free(ctx)    # the clone keeps the weights alive
    got = np.zeros(v, np.float32)
    lib.rwkv_eval(clone, HELLO[0], None, state.ctypes.data, got.ctypes.data)
    for t in HELLO[1:]:
        lib.rwkv_eval(clone, t, state.ctypes.data, state.ctypes.data, got.ctypes.data)
    assert got.tobytes() == want.tobytes()
    lib.rwkv_free(clone)


def test_init_state_equals_null_state(lib):
    """rwkv.h:197: rwkv_init_state(state) then eval == eval with state_in NULL; zero state is different for v4."""
    for ver in ("4v0-660K", "6v0-3m"):
        ctx = lib.rwkv_init_from_file(model_path(ver, "FP32"), 1, 0)
        n, v = lib.rwkv_get_state_len(ctx), lib.rwkv_get_logits_len(ctx)
        init = np.ones(n, np.float32)
        lib.rwkv_init_state(ctx, init.ctypes.data)
        a, b, sa, sb = np.zeros(v, np.float32), np.zeros(v, np.float32), np.zeros(n, np.float32), np.zeros(n, np.float32)
        lib.rwkv_eval(ctx, 65, None, sa.ctypes.data, a.ctypes.data)
        lib.rwkv_eval(ctx, 65, init.ctypes.data, sb.ctypes.data, b.ctypes.data)
        assert a.tobytes() == b.tobytes() and sa.tobytes() == sb.tobytes()
        lib.rwkv_free(ctx)


def test_argument_errors(lib):
    """rwkv_eval.inc:43,89,108,167-168: bad arguments return false and set RWKV_ERROR_ARGS on the context."""
    ctx = lib.rwkv_init_from_file(model_path("4v0-660K", "FP32"), 1, 0)
    assert lib.rwkv_get_last_error(None) == 0           # tests/logit_difference_validator.inc:35-36
    lib.rwkv_set_print_errors(ctx, False)
    n, v = lib.rwkv_get_state_len(ctx), lib.rwkv_get_logits_len(ctx)
    state, logits = np.zeros(n, np.float32), np.zeros(v, np.float32)
    P = ctypes.POINTER(ctypes.c_float)
    sp, lp = state.ctypes.data_as(P), logits.ctypes.data_as(P)
    assert not lib.library.rwkv_eval(ctx.ptr, 256, None, sp, lp)
    assert lib.rwkv_get_last_error(ctx) == 1 << 8 and lib.rwkv_get_last_error(ctx) == 0
    toks = (ctypes.c_uint32 * 3)(1, 2, 999)
    assert not lib.library.rwkv_eval_sequence(ctx.ptr, toks, 3, None, sp, lp)
    assert lib.rwkv_get_last_error(ctx) == 1 << 8
    assert not lib.library.rwkv_eval_sequence(ctx.ptr, toks, 0, None, sp, lp)
    assert not lib.library.rwkv_eval_sequence_in_chunks(ctx.ptr, toks, 2, 0, None, sp, lp)
    assert lib.library.rwkv_eval_sequence(ctx.ptr, None, 5, None, sp, lp)       # NULL tokens = warm-up only (rwkv.h:135)
    assert lib.library.rwkv_eval(ctx.ptr, 5, None, None, None)                  # every output optional
    assert lib.rwkv_get_n_vocab(ctx) == 256 and lib.rwkv_get_n_embed(ctx) == 128 and lib.rwkv_get_n_layer(ctx) == 4
    assert lib.rwkv_get_state_buffer_element_count(ctx) == n == 128 * 5 * 4
    lib.rwkv_free(ctx)
    lib.library.rwkv_free(None)                                                 # NULL-safe (rwkv.cpp:188)
    assert "B200" in lib.rwkv_get_system_info_string() or "DEVICE" in lib.rwkv_get_system_info_string()


def test_long_sequence_crosses_internal_pass_limit(models):
    """Sequences longer than the engine's per-pass token cap (512) are cut internally; result must equal chunked eval."""
    m = models(model_path("6v0-3m", "FP32"))
    toks = [(7919 * i + 3) % 256 for i in range(1100)]
    a_logits, a_state = m.eval_sequence(toks, None, use_numpy=True)
    b_logits, b_state = m.eval_sequence_in_chunks(toks, None, chunk_size=37, use_numpy=True)
    assert a_state.tobytes() == b_state.tobytes() and a_logits.tobytes() == b_logits.tobytes()
    assert np.isfinite(a_logits).all()


@pytest.mark.parametrize("preset,fmt,tol", [
    ("rwkv4-small", "FP32", 2e-4), ("rwkv5.1-small", "FP32", 2e-4), ("rwkv5-small", "FP32", 2e-4), ("rwkv6-small", "FP32", 2e-4), ("rwkv7-small", "FP32", 2e-4),
    ("rwkv6-small", "FP16", 2e-2), ("rwkv6-small", "Q5_1", 2e-1), ("rwkv7-small", "Q8_0", 2e-1), ("rwkv5-small", "Q4_0", 2e-1), ("rwkv4-small", "Q4_1", 2e-1),
    ("rwkv5.1-small", "Q5_0", 2e-1),
])
def test_synthetic_real_head_size_vs_oracle(pkg, lib, tmp_path, preset, fmt, tol):
    """Head size 64 / LoRA ranks / FFN widths of real checkpoints (the fixtures only have head size 8): synthetic weights,
    ours vs the numpy oracle. FP32 is tight; FP16 / quantised bars allow for rounding-flip divergence over 8 tokens."""
    import rwkv_oracle as ro
    import synthetic_model as sm
    path = str(tmp_path / f"{preset}-{fmt}.bin")
    sm.write_direct(path, preset, fmt, seed=11)
    oracle = ro.OracleModel(path)
    toks = sm.synthetic_tokens(8, oracle.n_vocab)
    want_logits, want_state = oracle.eval_sequence(toks)
    m = pkg.RWKVModel(lib, path, thread_count=1)
    logits, state = serial(m, toks)
    seq_logits, seq_state = m.eval_sequence(toks, None, use_numpy=True)
    m.free()
    assert np.array_equal(seq_logits, logits) and np.array_equal(seq_state, state)
    assert np.abs(logits - want_logits).max() <= tol, np.abs(logits - want_logits).max()
    assert np.abs(state - want_state).max() <= 50 * tol


def test_reference_c_tests_run_against_our_library(lib, tmp_path):
    """The reference's own C test programs (tests/*.c), compiled UNMODIFIED against include/rwkv.h and linked to OUR
    librwkv.so by `make -C oracle ctests` (binaries in oracle/_ref/ctests/, built where /root/reference exists)."""
    bindir = os.path.join(ROOT, "oracle", "_ref", "ctests")
    if not os.path.isdir(bindir) or not os.listdir(bindir):
        pytest.skip("oracle/_ref/ctests not built")
    work = tmp_path / "ctests"
    work.mkdir()
    for f in os.listdir(os.path.join(ROOT, "tests", "golden", "models")):
        os.symlink(os.path.join(ROOT, "tests", "golden", "models", f), work / f)
    for f in os.listdir(os.path.join(ROOT, "tests", "golden", "logits")):
        os.symlink(os.path.join(ROOT, "tests", "golden", "logits", f), work / f)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(lib.path) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    for exe in sorted(os.listdir(bindir)):
        r = subprocess.run([os.path.join(bindir, exe)], cwd=work, env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (exe, r.stderr[-2000:])


def test_two_clones_on_two_host_threads(lib):
    """rwkv.h:65-67, 94-96: one evaluation at a time per context, but clones may run concurrently on different threads. Two clones
    evaluate different token streams from two host threads (ctypes releases the GIL), one of them with the in-kernel timeline
    switched on; each must produce exactly what it produces alone."""
    import threading
    path = model_path("6v0-3m", "Q5_1")
    ctx = lib.rwkv_init_from_file(path, 1, 0)
    clone = lib.rwkv_clone_context(ctx, 1)
    n, v = lib.rwkv_get_state_len(ctx), lib.rwkv_get_logits_len(ctx)
    P = ctypes.POINTER(ctypes.c_float)
    streams = {0: [(31 * i + 7) % 256 for i in range(150)], 1: [(17 * i + 3) % 256 for i in range(150)]}

    def run(c, toks, out):
        state, logits = np.zeros(n, np.float32), np.zeros(v, np.float32)
        for i, t in enumerate(toks):
            assert lib.library.rwkv_eval(c.ptr, t, None if i == 0 else state.ctypes.data_as(P), state.ctypes.data_as(P), logits.ctypes.data_as(P))
        out.append((logits.copy(), state.copy()))

    alone = {0: [], 1: []}
    run(ctx, streams[0], alone[0])
    run(clone, streams[1], alone[1])
    assert lib.library.rwkv_b200_trace_enable(clone.ptr)
    for _ in range(3):
        together = {0: [], 1: []}
        th = [threading.Thread(target=run, args=(ctx, streams[0], together[0])), threading.Thread(target=run, args=(clone, streams[1], together[1]))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for k in (0, 1):
            assert together[k] and together[k][0][0].tobytes() == alone[k][0][0].tobytes() and together[k][0][1].tobytes() == alone[k][0][1].tobytes(), k
    lib.library.rwkv_b200_trace_disable(clone.ptr)
    lib.rwkv_free(clone)
    lib.rwkv_free(ctx)
