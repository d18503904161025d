"""Workload for ncu: a few single-token decode steps (CUDA graphs off so every kernel is a separate launch) and
optionally one prefill chunk, on a synthetic model. Never a source of benchmark numbers.

    ncu ... python tools/profile_decode.py rwkv6-7b:Q5_1 --tokens 3 [--prefill 128]
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__  # noqa: E402
import bench  # noqa: E402
import synthetic_model as sm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("workload")
ap.add_argument("--tokens", type=int, default=3)
ap.add_argument("--prefill", type=int, default=0)
ap.add_argument("--graphs", action="store_true")
a = ap.parse_args()
pkg = __graft_entry__.load_package()
lib = pkg.load_rwkv_shared_library()
path, preset = bench.workload_file(a.workload)
ctx = lib.rwkv_b200_init_from_file_ex(path, 0, 0, -1)
lib.library.rwkv_b200_set_graphs(ctx.ptr, a.graphs)
toks = sm.synthetic_tokens(max(a.tokens, a.prefill) + 1, preset["V"])
arr = (ctypes.c_uint32 * len(toks))(*toks)
lib.library.rwkv_b200_state_load(ctx.ptr, None)
for i in range(a.tokens):
    lib.library.rwkv_b200_eval_resident(ctx.ptr, ctypes.cast(ctypes.byref(arr, 4 * i), ctypes.POINTER(ctypes.c_uint32)), 1, True, None)
if a.prefill:
    lib.library.rwkv_b200_eval_resident(ctx.ptr, arr, a.prefill, True, None)
lib.library.rwkv_b200_synchronize(ctx.ptr)
lib.rwkv_free(ctx)
print("profile workload done")
