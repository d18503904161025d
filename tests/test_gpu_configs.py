"""BASELINE.json's configurations at their FULL shapes (synthetic weights, tools/synthetic_model.py), checked through the
size-independent properties the reference's own tests pin (tests/test_eval_sequence_in_chunks.c, test_logit_calculation_skipping.c,
test_context_cloning.c): serial == sequence == chunked bit for bit on the batch-invariant path, logits on / off leave the state
alone, a clone continues identically; plus the engine's own invariants at these sizes: overlapped state copies == plain copies,
and the tensor-core prefill path stays close to the batch-invariant one (the persistent kernel is checked out of process in
tests/test_zz_gpu_persistent.py)."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P_F = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def synth(tmp_path_factory):
    import synthetic_model as sm
    d = tmp_path_factory.mktemp("configs")
    made = {}

    def get(preset, fmt):
        key = (preset, fmt)
        if key not in made:
            path = str(d / f"{preset}-{fmt}.bin")
            sm.write_direct(path, preset, fmt, seed=3)
            made[key] = path
        return made[key]
    return get


def eval_serial(lib, ctx, toks, n_state, n_logits, logits_every=1):
    state = np.zeros(n_state, np.float32)
    logits = np.zeros(n_logits, np.float32)
    for i, t in enumerate(toks):
        want = (i + 1) % logits_every == 0 or i == len(toks) - 1
        assert lib.library.rwkv_eval(ctx.ptr, t, None if i == 0 else state.ctypes.data_as(P_F), state.ctypes.data_as(P_F), logits.ctypes.data_as(P_F) if want else None)
    return logits, state


@pytest.mark.parametrize("preset,fmt", [
    ("rwkv4-169m", "Q5_1"), ("rwkv5-1b5", "Q4_0"),       # BASELINE.json configs 1 and 2: green on B200 in round 1
    # config 3 (5.8 GB synthetic file) was added after round 1's last GPU run: until it has been seen green it only runs in a child
    # process (tests/test_zz_gpu_persistent.py sets RWKV_RUN_UNVALIDATED=1), where a fault cannot take the suite's CUDA context along
    pytest.param("rwkv7-2b9", "FP16", marks=pytest.mark.skipif(os.environ.get("RWKV_RUN_UNVALIDATED") != "1", reason="not yet validated on a GPU; runs out of process")),
])
def test_full_shape_invariants(lib, synth, preset, fmt):
    import synthetic_model as sm
    ctx = lib.rwkv_init_from_file(synth(preset, fmt), 1, 0)
    try:
        n_state, n_logits = lib.rwkv_get_state_buffer_element_count(ctx), lib.rwkv_get_logits_buffer_element_count(ctx)
        toks = sm.synthetic_tokens(24, n_logits)
        arr = (ctypes.c_uint32 * len(toks))(*toks)
        lib.library.rwkv_b200_set_persistent(ctx.ptr, False)
        lib.library.rwkv_b200_set_overlap(ctx.ptr, False)
        logits, state = eval_serial(lib, ctx, toks, n_state, n_logits)
        assert np.isfinite(logits).all() and np.isfinite(state).all() and logits.std() > 0
        # sequence mode and chunked mode (chunks below the tensor-core threshold of 32 tokens): bit-identical to serial
        s2, l2 = np.zeros(n_state, np.float32), np.zeros(n_logits, np.float32)
        assert lib.library.rwkv_eval_sequence(ctx.ptr, arr, len(toks), None, s2.ctypes.data_as(P_F), l2.ctypes.data_as(P_F))
        assert s2.tobytes() == state.tobytes() and l2.tobytes() == logits.tobytes()
        for chunk in (1, 5, 10):
            s3, l3 = np.zeros(n_state, np.float32), np.zeros(n_logits, np.float32)
            assert lib.library.rwkv_eval_sequence_in_chunks(ctx.ptr, arr, len(toks), chunk, None, s3.ctypes.data_as(P_F), l3.ctypes.data_as(P_F))
            assert s3.tobytes() == state.tobytes() and l3.tobytes() == logits.tobytes(), chunk
        # logits only every 4th token: same state, same final logits
        l4, s4 = eval_serial(lib, ctx, toks, n_state, n_logits, logits_every=4)
        assert s4.tobytes() == state.tobytes() and l4.tobytes() == logits.tobytes()
        # a clone picks the sequence up in the middle
        l5, s5 = eval_serial(lib, ctx, toks[:10], n_state, n_logits)
        clone = lib.rwkv_clone_context(ctx, 1)
        for t in toks[10:]:
            assert lib.library.rwkv_eval(clone.ptr, t, s5.ctypes.data_as(P_F), s5.ctypes.data_as(P_F), l5.ctypes.data_as(P_F))
        lib.rwkv_free(clone)
        assert s5.tobytes() == state.tobytes() and l5.tobytes() == logits.tobytes()
        # engine invariant: the host state copied per layer group, overlapped with the kernels (the default of rwkv_eval)
        lib.library.rwkv_b200_set_overlap(ctx.ptr, True)
        l6, s6 = eval_serial(lib, ctx, toks, n_state, n_logits)
        assert s6.tobytes() == state.tobytes() and l6.tobytes() == logits.tobytes()
    finally:
        lib.rwkv_free(ctx)


@pytest.mark.skipif(os.environ.get("RWKV_RUN_UNVALIDATED") != "1", reason="not yet validated on a GPU (round 1's last run lost its CUDA context before reaching it); runs out of process")
def test_prefill_chunk_128_tensor_core_path_close_to_serial(lib, synth):
    """Config 2 of BASELINE.json: RWKV-5-World-1.5B shape, Q4_0, one 128-token chunk. The tcgen05 path multiplies fp16-rounded
    activations (the reference multiplies int8-quantised ones), so it is compared with the batch-invariant path with a tolerance;
    with the tensor cores switched off the chunk must equal serial evaluation bit for bit."""
    import synthetic_model as sm
    ctx = lib.rwkv_init_from_file(synth("rwkv5-1b5", "Q4_0"), 1, 0)
    try:
        n_state, n_logits = lib.rwkv_get_state_buffer_element_count(ctx), lib.rwkv_get_logits_buffer_element_count(ctx)
        toks = sm.synthetic_tokens(128, n_logits)
        arr = (ctypes.c_uint32 * 128)(*toks)
        lib.library.rwkv_b200_set_tensor_cores(ctx.ptr, False)
        s0, l0 = np.zeros(n_state, np.float32), np.zeros(n_logits, np.float32)
        assert lib.library.rwkv_eval_sequence_in_chunks(ctx.ptr, arr, 128, 128, None, s0.ctypes.data_as(P_F), l0.ctypes.data_as(P_F))
        l1, s1 = eval_serial(lib, ctx, toks, n_state, n_logits, logits_every=1000)
        assert s0.tobytes() == s1.tobytes() and l0.tobytes() == l1.tobytes()
        lib.library.rwkv_b200_set_tensor_cores(ctx.ptr, True)
        s2, l2 = np.zeros(n_state, np.float32), np.zeros(n_logits, np.float32)
        assert lib.library.rwkv_eval_sequence_in_chunks(ctx.ptr, arr, 128, 128, None, s2.ctypes.data_as(P_F), l2.ctypes.data_as(P_F))
        assert np.isfinite(l2).all() and np.isfinite(s2).all()
        rel = float(np.linalg.norm(l2 - l0) / np.linalg.norm(l0))
        assert rel <= 0.25, rel          # measured on B200: see DESIGN.md section 3 (activation rounding differs, fp16 vs int8 blocks)
    finally:
        lib.rwkv_free(ctx)
