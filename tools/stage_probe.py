"""Experiment: where do 12 % of the single-GPU decode step go?  Times (CUDA events on the default stream, graphs on)
  A  the whole model, one context                      (= bench.py decode)
  B  the whole model, two cloned contexts alternating  (what a pipeline stage does)
  C  layers [0, L/2) only, one context / two contexts alternating
Never a source of benchmark numbers."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import __graft_entry__, bench
import synthetic_model as sm
pkg = __graft_entry__.load_package(); lib = pkg.load_rwkv_shared_library(); L = lib.library
path, preset = bench.workload_file(sys.argv[1] if len(sys.argv) > 1 else "rwkv6-7b:Q5_1")
PU = ctypes.POINTER(ctypes.c_uint32)
toks = sm.synthetic_tokens(600, preset["V"]); arr = (ctypes.c_uint32 * 600)(*toks)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); sp = ctypes.c_void_p(stream.cuda_stream)   # handle 0 (legacy default stream) would mean "the context's own stream"


def run(ctxs, steps, hidden_out, own_stream=False):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    def one(i):
        c = ctxs[i % len(ctxs)]
        assert L.rwkv_b200_stage_eval(c.ptr, ctypes.cast(ctypes.byref(arr, 4 * i), PU), 1, None,
                                      ctypes.c_void_p(hidden_out.data_ptr()) if hidden_out is not None else None, True, None if own_stream else sp)
    for i in range(16): one(i)
    torch.cuda.synchronize()
    for c in ctxs: L.rwkv_b200_synchronize(c.ptr)
    if own_stream:
        import time
        t0 = time.perf_counter()
        for i in range(steps): one(16 + i)
        for c in ctxs: L.rwkv_b200_synchronize(c.ptr)
        return (time.perf_counter() - t0) * 1e3 / steps
    e0.record(stream)
    for i in range(steps): one(16 + i)
    e1.record(stream); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


nl = preset["L"]
full = lib.rwkv_b200_init_from_file_ex(path, 0, 0, -1)
full2 = lib.rwkv_clone_context(full, 1)
for c in (full, full2): L.rwkv_b200_state_load(c.ptr, None)
print("A  whole model, 1 context, torch stream : %.3f ms/token" % run([full], 200, None))
print("A' whole model, 1 context, own stream   : %.3f ms/token (host clock)" % run([full], 200, None, True))
print("B' whole model, 2 contexts, own streams : %.3f ms/token (host clock; the two graphs overlap on the GPU)" % run([full, full2], 200, None, True))
print("B  whole model, 2 contexts alternating  : %.3f ms/token" % run([full, full2], 200, None))
lib.rwkv_free(full2); lib.rwkv_free(full)
half = lib.rwkv_b200_init_from_file_ex(path, 0, 0, nl // 2)
half2 = lib.rwkv_clone_context(half, 1)
for c in (half, half2): L.rwkv_b200_state_load(c.ptr, None)
buf = torch.zeros(L.rwkv_b200_stage_hidden_len(half.ptr, 1), dtype=torch.float32, device="cuda:0")
print("C  layers [0, %d), 1 context             : %.3f ms/token" % (nl // 2, run([half], 200, buf)))
print("C' layers [0, %d), 2 contexts alternating: %.3f ms/token" % (nl // 2, run([half, half2], 200, buf)))

# ---- which kernels are faster when two contexts alternate? in-kernel timeline of context 0 in both modes
def timeline(ctxs, label):
    c0 = ctxs[0]
    assert L.rwkv_b200_trace_enable(c0.ptr)
    N = 1024
    st = (ctypes.c_double * N)(); en = (ctypes.c_double * N)(); names = ctypes.create_string_buffer(32 * N)
    for i in range(12):
        c = ctxs[i % len(ctxs)]
        assert L.rwkv_b200_stage_eval(c.ptr, ctypes.cast(ctypes.byref(arr, 4 * i), PU), 1, None, None, True, sp)
        if c is c0:
            torch.cuda.synchronize()
            n = L.rwkv_b200_trace_read(c0.ptr, st, en, ctypes.cast(names, ctypes.c_void_p), N)
    rows = [(names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(), st[i], en[i]) for i in range(n)]
    att = {}; pe = None
    for nm, s_, e_ in rows:
        if pe is None: pe = s_
        att.setdefault(nm, [0, 0.0]); att[nm][0] += 1
        if e_ > pe: att[nm][1] += e_ - pe; pe = e_
    print(label, "span %.1f us:" % max(r[2] for r in rows), {k: (v[0], round(v[1], 1)) for k, v in att.items()})
    first = rows[:6]
    print("   first kernels:", [(nm, round(s_, 1), round(e_ - s_, 1)) for nm, s_, e_ in first])

lib.rwkv_free(half2); lib.rwkv_free(half)
full = lib.rwkv_b200_init_from_file_ex(path, 0, 0, -1)
full2 = lib.rwkv_clone_context(full, 1)
for c in (full, full2): L.rwkv_b200_state_load(c.ptr, None)
timeline([full], "solo       ")
timeline([full, full2], "alternating")
