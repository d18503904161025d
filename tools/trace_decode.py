"""GPU-side timeline of one decode step in CUDA-graph mode (our nsys substitute): prints per-kernel start/duration/gap.
    python tools/trace_decode.py rwkv6-7b:Q5_1 [--layers 2] [--out gpurun_out/trace.csv]"""
import argparse, ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__, bench
import synthetic_model as sm
ap = argparse.ArgumentParser(); ap.add_argument("workload"); ap.add_argument("--out", default=None); ap.add_argument("--show", type=int, default=40); ap.add_argument("--prefill", type=int, default=0)
a = ap.parse_args()
pkg = __graft_entry__.load_package(); lib = pkg.load_rwkv_shared_library(); L = lib.library
path, preset = bench.workload_file(a.workload)
ctx = lib.rwkv_b200_init_from_file_ex(path, 0, 0, -1)
toks = sm.synthetic_tokens(64, preset["V"]); arr = (ctypes.c_uint32 * 64)(*toks)
PU = ctypes.POINTER(ctypes.c_uint32)
L.rwkv_b200_state_load(ctx.ptr, None)
assert L.rwkv_b200_trace_enable(ctx.ptr)
N = 1024
st = (ctypes.c_double * N)(); en = (ctypes.c_double * N)(); names = ctypes.create_string_buffer(32 * N)
marks = (ctypes.c_double * (4 * N))()
L.rwkv_b200_trace_set_marks_buffer.argtypes = [ctypes.POINTER(ctypes.c_double)]
L.rwkv_b200_trace_set_marks_buffer(marks)
if a.prefill:
    toks = sm.synthetic_tokens(a.prefill, preset["V"]); arr = (ctypes.c_uint32 * a.prefill)(*toks)
    for i in range(3):
        L.rwkv_b200_eval_resident(ctx.ptr, arr, a.prefill, True, None)
        n = L.rwkv_b200_trace_read(ctx.ptr, st, en, ctypes.cast(names, ctypes.c_void_p), N)
else:
    for i in range(8):   # eager, eager, capture, replay...
        L.rwkv_b200_eval_resident(ctx.ptr, ctypes.cast(ctypes.byref(arr, 4 * i), PU), 1, True, None)
        n = L.rwkv_b200_trace_read(ctx.ptr, st, en, ctypes.cast(names, ctypes.c_void_p), N)
rows = [(names.raw[32 * i:32 * i + 32].split(b"\0")[0].decode(), st[i], en[i]) for i in range(n)]
total = max(r[2] for r in rows)
print("records", n, "step span %.1f us" % total)
agg = {}
prev_end = 0.0
lines = []
for i, (nm, s, e) in enumerate(rows):
    gap = s - prev_end
    agg.setdefault(nm, [0, 0.0, 0.0]); agg[nm][0] += 1; agg[nm][1] += e - s; agg[nm][2] += max(gap, 0)
    lines.append("%4d %-18s start %9.1f dur %7.1f gap %6.1f" % (i, nm, s, e - s, gap))
    prev_end = max(prev_end, e)
for l in lines[:a.show]: print(l)
print("...")
for l in lines[-6:]: print(l)
for nm, (c, d, g) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-18s n=%4d busy %8.1f us  (avg %6.2f)  gaps-before %7.1f us" % (nm, c, d, d / c, g))
# critical-path attribution: each kernel is charged the time by which it pushed the end of the timeline out
att = {}; pe = None
for nm, s_, e_ in rows:
    if pe is None: pe = s_
    att.setdefault(nm, [0, 0.0]); att[nm][0] += 1
    if e_ > pe: att[nm][1] += e_ - pe; pe = e_
print("critical-path attribution:")
for nm, (c, d) in sorted(att.items(), key=lambda x: -x[1][1]): print("  %-16s n=%4d  %9.1f us  (avg %7.2f)" % (nm, c, d, d / c))
# one layer from the middle, with the in-kernel cycle counters of the kernels that report them
mid = [i for i, r in enumerate(rows) if r[0].startswith("wkv")]
if mid:
    k = mid[len(mid) // 2]; pe = None
    for i in range(max(0, k - 14), min(len(rows), k + 8)):
        nm, s_, e_ = rows[i]
        cyc = [int(round((marks[4 * i + q] - s_) * 1000)) for q in range(4)] if marks[4 * i] >= 0 else ""
        print("  %-12s start %9.1f dur %7.1f crit %7.1f %s" % (nm, s_, e_ - s_, e_ - (pe if pe is not None else s_), cyc))
        pe = max(pe or 0, e_)
try:
    import pynvml
    pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0); clk = []
    for i in range(10):
        L.rwkv_b200_eval_resident(ctx.ptr, arr, a.prefill or 1, True, None)
        clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
    print("sm clock MHz under load:", clk)
except Exception as ex:
    print("clock sample failed:", ex)
if a.out:
    with open(a.out, "w") as f:
        f.write("idx,name,start_us,end_us,m0_after_wait,m1_after_stage,m2_after_tile0,m3_after_tile2\n")
        for i, (nm, s, e) in enumerate(rows): f.write(f"{i},{nm},{s:.3f},{e:.3f}," + ",".join("%.3f" % marks[4 * i + k] for k in range(4)) + "\n")
lib.rwkv_free(ctx)
