"""Aggregate decode throughput of batch contexts (rwkv_b200_batch_*) at the 7B shape: tokens/s for B = 1 (plain context) .. 64.
Measurement aid, not a bench line."""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__, bench
import synthetic_model as sm
pkg = __graft_entry__.load_package(); lib = pkg.load_rwkv_shared_library(); L = lib.library
spec = sys.argv[1] if len(sys.argv) > 1 else "rwkv6-7b:Q5_1"
path, preset = bench.workload_file(spec)
ctx = lib.rwkv_b200_init_from_file_ex(path, 0, 0, -1)
V = preset["V"]
toks = sm.synthetic_tokens(64, V); arr = (ctypes.c_uint32 * 64)(*toks)
L.rwkv_b200_state_load(ctx.ptr, None)
ms = L.rwkv_b200_time_resident(ctx.ptr, arr, 1, 32, 8, True)
print(f"{spec}: single sequence {ms / 32:.3f} ms/token = {32e3 / ms:.0f} tok/s")
for B in (2, 4, 8, 16, 32, 64):
    b = ctypes.c_void_p(L.rwkv_b200_batch_create(ctx.ptr, B))
    if not b:
        print("batch_create failed for", B); continue
    t = (ctypes.c_uint32 * B)(*[(7919 * i) % V for i in range(B)])
    for _ in range(4):
        assert L.rwkv_b200_batch_eval(b, t, True)
    L.rwkv_b200_synchronize(b)
    n = 12
    t0 = time.perf_counter()
    for _ in range(n):
        assert L.rwkv_b200_batch_eval(b, t, True)
    L.rwkv_b200_synchronize(b)
    dt = (time.perf_counter() - t0) / n
    print(f"  B = {B:3d}: {dt * 1e3:7.3f} ms per step = {B / dt:8.0f} tok/s aggregate ({'tensor cores' if B >= 16 else 'multi-column GEMV'})")
    L.rwkv_free(b)
lib.rwkv_free(ctx)
