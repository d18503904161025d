"""Pipeline-stage entry points (include/rwkv_b200.h) on ONE GPU: a model cut into two stage contexts, chained through a
device buffer, must give exactly the logits of the whole model -- token by token (serial path, CUDA graphs) and for one sequence pass."""
import ctypes

import numpy as np
import pytest

from conftest import LONG_PROMPT, VERSIONS, model_path

pytestmark = pytest.mark.gpu
PF = ctypes.POINTER(ctypes.c_float)
PU = ctypes.POINTER(ctypes.c_uint32)


@pytest.mark.parametrize("ver", VERSIONS)
@pytest.mark.parametrize("fmt", ["FP32", "Q5_1"])
def test_two_stages_on_one_gpu_match_the_whole_model(pkg, lib, ver, fmt):
    import torch
    L = lib.library
    path = model_path(ver, fmt)
    whole = pkg.RWKVModel(lib, path, thread_count=1)
    n_layer, n_vocab = lib.rwkv_get_n_layer(whole._ctx), lib.rwkv_get_logits_len(whole._ctx)
    cut = pkg.pipeline.stage_layers(n_layer, 2, 0)[1]
    a = lib.rwkv_b200_init_from_file_ex(path, 0, 0, cut)
    b = lib.rwkv_b200_init_from_file_ex(path, 0, cut, -1)
    try:
        for c in (a, b):
            assert L.rwkv_b200_state_load(c.ptr, None)
        tokens = LONG_PROMPT[:12]
        n_hidden = L.rwkv_b200_stage_hidden_len(a.ptr, 1)
        assert n_hidden == pkg.pipeline.hidden_floats(lib.rwkv_get_n_embed(whole._ctx), 1, 7 if ver.startswith("7") else 6)
        buf = torch.zeros(n_hidden, dtype=torch.float32, device="cuda:0")
        got = np.zeros(n_vocab, np.float32)
        state = None
        for t in tokens:       # serial: the third token on replays the captured graphs of both stages
            want, state = whole.eval(t, state, use_numpy=True)
            tok = (ctypes.c_uint32 * 1)(t)
            assert L.rwkv_b200_stage_eval(a.ptr, tok, 1, None, ctypes.c_void_p(buf.data_ptr()), False, None)
            assert L.rwkv_b200_synchronize(a.ptr)
            assert L.rwkv_b200_stage_eval(b.ptr, None, 1, ctypes.c_void_p(buf.data_ptr()), None, True, None)
            assert L.rwkv_b200_stage_logits(b.ptr, got.ctypes.data_as(PF), None)
            assert np.array_equal(got, want), (ver, fmt, t)
        # one sequence pass of 9 tokens continuing from the same states
        seq = LONG_PROMPT[12:21]
        want, state = whole.eval_sequence(seq, state, use_numpy=True)
        arr = (ctypes.c_uint32 * len(seq))(*seq)
        buf = torch.zeros(L.rwkv_b200_stage_hidden_len(a.ptr, len(seq)), dtype=torch.float32, device="cuda:0")
        assert L.rwkv_b200_stage_eval(a.ptr, arr, len(seq), None, ctypes.c_void_p(buf.data_ptr()), False, None)
        assert L.rwkv_b200_synchronize(a.ptr)
        assert L.rwkv_b200_stage_eval(b.ptr, None, len(seq), ctypes.c_void_p(buf.data_ptr()), None, True, None)
        assert L.rwkv_b200_stage_logits(b.ptr, got.ctypes.data_as(PF), None)
        assert np.array_equal(got, want)
        # argument checks: a later stage needs activations, an earlier one an output buffer
        assert not L.rwkv_b200_stage_eval(b.ptr, arr, 1, None, None, True, None)
        assert not L.rwkv_b200_stage_eval(a.ptr, arr, 1, None, None, False, None)
    finally:
        lib.rwkv_free(a); lib.rwkv_free(b); whole.free()


@pytest.mark.parametrize("ver,fmt", [("6v0-3m", "Q5_1"), ("7v0-834K", "FP16"), ("4v0-660K", "FP32"), ("5v2-730K", "Q5_1")])
def test_peer_memory_pipeline_three_stages_two_sequences(pkg, lib, ver, fmt):
    """The hand-off inside the library (csrc/kernels/pipe.cu: mailbox + credit flags written with peer stores, device-side item
    counters): three stages on ONE GPU connected with rwkv_b200_pipe_connect_local, two sequences in flight (one context per
    sequence and stage), every stage on its own stream, the host never waiting for a neighbour. Logits must equal the whole
    model's bit for bit, token by token (graph replays) and for multi-token passes; v7 carries v_first through the mailbox."""
    L = lib.library
    path = model_path(ver, fmt)
    whole = [pkg.RWKVModel(lib, path, thread_count=1) for _ in range(2)]
    n_layer, n_vocab = lib.rwkv_get_n_layer(whole[0]._ctx), lib.rwkv_get_logits_len(whole[0]._ctx)
    cuts = [pkg.pipeline.stage_layers(n_layer, 3, r) for r in range(3)]
    stages = [lib.rwkv_b200_init_from_file_ex(path, 0, b, e) for b, e in cuts]
    seqs = [[s] + [lib.rwkv_clone_context(s, 1)] for s in stages]          # seqs[stage][sequence]
    try:
        for r in range(3):
            assert L.rwkv_b200_pipe_connect_local(stages[r].ptr, stages[r - 1].ptr if r > 0 else None, stages[r + 1].ptr if r < 2 else None)
            for c in seqs[r]:
                assert L.rwkv_b200_state_load(c.ptr, None)
        sp = [ctypes.c_void_p(L.rwkv_b200_stream(stages[r].ptr)) for r in range(3)]      # ONE stream per stage: a link's items are enqueued in order
        streams = {0: LONG_PROMPT[:14], 1: [(5 * i + 11) % 256 for i in range(14)]}
        states = [None, None]
        got = np.zeros(n_vocab, np.float32)
        for i in range(14):                     # one token per pass; the third pass on replays the captured graphs
            for q in (0, 1):
                tok = (ctypes.c_uint32 * 1)(streams[q][i])
                for r in range(3):
                    assert L.rwkv_b200_pipe_eval(seqs[r][q].ptr, tok if r == 0 else None, 1, True, sp[r]), (r, q, i)
            for q in (0, 1):
                want, states[q] = whole[q].eval(streams[q][i], states[q], use_numpy=True)
                # the last stage evaluated sequence 0 then 1 on ONE context each: read each one's logits
                assert L.rwkv_b200_stage_logits(seqs[2][q].ptr, got.ctypes.data_as(PF), sp[2])
                assert np.array_equal(got, want), (ver, fmt, i, q)
        for q in (0, 1):                        # a 9-token pass and a 40-token pass (tensor-core path for non-F32 weights)
            for n in (9, 40):
                seq = [(7 * j + 3 * q + n) % 256 for j in range(n)]
                arr = (ctypes.c_uint32 * n)(*seq)
                for r in range(3):
                    assert L.rwkv_b200_pipe_eval(seqs[r][q].ptr, arr if r == 0 else None, n, True, sp[r])
                want, states[q] = whole[q].eval_sequence(seq, states[q], use_numpy=True)
                assert L.rwkv_b200_stage_logits(seqs[2][q].ptr, got.ctypes.data_as(PF), sp[2])
                if n < 32 or fmt == "FP32":
                    assert np.array_equal(got, want), (ver, fmt, n, q)
                else:       # K-splits of the tensor-core GEMM depend on the stage's matrix set only through the launch shape: same bits expected too
                    assert np.abs(got - want).max() <= 5e-2, (ver, fmt, n, q, np.abs(got - want).max())
        assert not L.rwkv_b200_pipe_eval(seqs[0][0].ptr, None, 1, False, None)          # the first stage needs tokens
    finally:
        for r in range(3):
            lib.rwkv_free(seqs[r][1]); lib.rwkv_free(seqs[r][0])
        for w in whole:
            w.free()


@pytest.mark.parametrize("ver,fmt", [("6v0-3m", "Q5_1"), ("7v0-834K", "FP32"), ("5v1-730K", "FP16")])
def test_in_process_pipeline_behind_rwkv_h(pkg, lib, ver, fmt):
    """rwkv_b200_init_pipeline / RWKV_B200_PIPELINE_DEVICES: several stages inside one process behind the PLAIN rwkv.h calls (here four
    stages on device 0; on a multi-GPU box one per device). rwkv_eval, rwkv_eval_sequence(_in_chunks), aliased / NULL buffers, clones
    evaluated from two host threads: everything must equal the single-context evaluation bit for bit (below the tensor-core
    threshold), and the stages' state slices must tile the caller's state buffer exactly."""
    import threading
    L = lib.library
    path = model_path(ver, fmt)
    one = pkg.RWKVModel(lib, path, thread_count=1)
    devs = (ctypes.c_int * 4)(0, 0, 0, 0)
    ptr = L.rwkv_b200_init_pipeline(path.encode(), devs, 4)
    assert ptr and L.rwkv_b200_pipeline_stages(ptr) == 4 and L.rwkv_b200_pipeline_stages(one._ctx.ptr) == 0
    pipe = ctypes.c_void_p(ptr)
    n, v = one.state_len, one.n_vocab
    try:
        assert lib.library.rwkv_get_state_len(pipe) == n and lib.library.rwkv_get_n_layer(pipe) == one.n_layer
        toks = LONG_PROMPT[:20]
        state, logits = np.zeros(n, np.float32), np.zeros(v, np.float32)
        want_state = None
        for i, t in enumerate(toks):             # serial, state_in aliases state_out from the second call on
            want, want_state = one.eval(t, want_state, use_numpy=True)
            assert L.rwkv_eval(pipe, t, None if i == 0 else state.ctypes.data_as(PF), state.ctypes.data_as(PF), logits.ctypes.data_as(PF))
            assert logits.tobytes() == want.tobytes() and state.tobytes() == want_state.tobytes(), (ver, fmt, i)
        arr = (ctypes.c_uint32 * len(toks))(*toks)
        s2, l2 = np.zeros(n, np.float32), np.zeros(v, np.float32)
        assert L.rwkv_eval_sequence(pipe, arr, len(toks), None, s2.ctypes.data_as(PF), l2.ctypes.data_as(PF))
        assert s2.tobytes() == want_state.tobytes() and l2.tobytes() == want.tobytes()
        s3, l3 = np.zeros(n, np.float32), np.zeros(v, np.float32)
        assert L.rwkv_eval_sequence_in_chunks(pipe, arr, len(toks), 7, None, s3.ctypes.data_as(PF), l3.ctypes.data_as(PF))
        assert s3.tobytes() == want_state.tobytes() and l3.tobytes() == want.tobytes()
        assert L.rwkv_eval(pipe, toks[0], None, None, None)                       # every output optional
        # two clones on two threads: their passes interleave on the stages' streams
        clones = [ctypes.c_void_p(L.rwkv_clone_context(pipe, 1)) for _ in range(2)]
        streams = [[(13 * i + 5 * q) % 256 for i in range(40)] for q in range(2)]
        wants = []
        for q in range(2):
            st = None
            for t in streams[q]:
                lg, st = one.eval(t, st, use_numpy=True)
            wants.append((lg.copy(), st.copy()))
        outs = [None, None]

        def run(q):
            st, lg = np.zeros(n, np.float32), np.zeros(v, np.float32)
            for i, t in enumerate(streams[q]):
                assert L.rwkv_eval(clones[q], t, None if i == 0 else st.ctypes.data_as(PF), st.ctypes.data_as(PF), lg.ctypes.data_as(PF))
            outs[q] = (lg, st)
        th = [threading.Thread(target=run, args=(q,)) for q in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for q in range(2):
            assert outs[q][0].tobytes() == wants[q][0].tobytes() and outs[q][1].tobytes() == wants[q][1].tobytes(), q
        for c in clones:
            L.rwkv_free(c)
    finally:
        L.rwkv_free(pipe)
        one.free()
