"""Workload for `ncu --set full` captures of EVERY kernel of the eval path (not only the GEMV / GEMM): a few passes over small
models that reach each kernel once or twice, CUDA graphs off so that ncu sees plain launches.

    ncu --set full --clock-control none --import-source on -o gpurun_out/r2_ncu_all python tools/ncu_targets.py
    python tools/ncu_summary.py gpurun_out/r2_ncu_all.ncu-rep profiles/r2_ncu_all_kernels_summary.txt

  rwkv6-wide  (ONE layer of the 7B shape, Q5_1)  decode token, 128-token chunk (tcgen05 path), 8-token chunk (multi-column GEMV,
                                                  tiled lerp), batch of 4 sequences, on-device sampling
  rwkv7-small (4 layers, FP16)                    decode token + 40-token chunk   -> wkv7_kernel, F16 GEMV / GEMM
  rwkv4-small (3 layers, Q4_1)                    decode token + 8-token chunk    -> wkv4_kernel
  rwkv5-small (3 layers, Q8_0)                    decode token                    -> wkv6_kernel with static decay
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import __graft_entry__  # noqa: E402
import synthetic_model as sm  # noqa: E402

PU = ctypes.POINTER(ctypes.c_uint32)


def main():
    pkg = __graft_entry__.load_package()
    lib = pkg.load_rwkv_shared_library()
    L = lib.library
    d = os.environ.get("RWKV_B200_BENCH_DIR", "/tmp/rwkv_b200_bench")
    os.makedirs(d, exist_ok=True)
    plan = [("rwkv6-wide", "Q5_1", [1, 128, 8], True), ("rwkv7-small", "FP16", [1, 40], False), ("rwkv4-small", "Q4_1", [1, 8], False), ("rwkv5-small", "Q8_0", [1], False)]
    for preset, fmt, passes, extras in plan:
        path = os.path.join(d, f"{preset}-{fmt}-seed5.bin")
        if not os.path.isfile(path):
            sm.write_direct(path, preset, fmt, seed=5)
        ctx = lib.rwkv_b200_init_from_file_ex(path, 0, 0, -1)
        L.rwkv_b200_set_graphs(ctx.ptr, False)
        n_vocab = lib.rwkv_get_logits_len(ctx)
        toks = sm.synthetic_tokens(256, n_vocab)
        arr = (ctypes.c_uint32 * len(toks))(*toks)
        L.rwkv_b200_state_load(ctx.ptr, None)
        for T in passes:
            assert L.rwkv_b200_eval_resident(ctx.ptr, arr, T, True, None)
        if extras:
            tok = ctypes.c_uint32(0)
            assert L.rwkv_b200_sample(ctx.ptr, ctypes.c_float(0.8), ctypes.c_float(0.5), ctypes.c_double(0.3), None, None, 0, ctypes.byref(tok))
            b = L.rwkv_b200_batch_create(ctx.ptr, 4)
            if b:
                assert L.rwkv_b200_batch_eval(b, arr, True)
                L.rwkv_b200_synchronize(b)
                L.rwkv_free(b)
        L.rwkv_b200_synchronize(ctx.ptr)
        lib.rwkv_free(ctx)
    print("ncu targets done")


if __name__ == "__main__":
    main()
