#!/usr/bin/env python
"""bench.py -- the north-star measurement: RWKV-6-World-7B-shape Q5_1, single-token decode and 128-token
prefill through the rwkv.h C ABI of rwkv.cpp_b200/librwkv.so, on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload rwkv6-7b:Q5_1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one single-token evaluation (with logits) of the whole model. Prints ONE JSON line:
  value        decode tokens/s, state resident in HBM, CUDA-event timed on the library's own stream
  e2e          the same metric through rwkv_eval() with HOST state/logits buffers (H2D + D2H inside the timed region)
  prefill      128-token rwkv_eval_sequence chunk: device-timed and e2e tokens/s
  roofline     dominant kernel (fused dequant-GEMV): algorithmic weight bytes / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline the UNMODIFIED reference (oracle/_ref) timed on this box's host cores on a bounded sample
Weights are synthetic (random-init, exact shapes/dtypes; tools/synthetic_model.py); no checkpoints exist offline.
The 6 GB of weights streamed per token exceed the 126 MB L2 48x, so no explicit L2 flush is needed between steps.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

PF = ctypes.POINTER(ctypes.c_float)
PU = ctypes.POINTER(ctypes.c_uint32)
CACHE_DIR = os.environ.get("RWKV_B200_BENCH_DIR", "/tmp/rwkv_b200_bench")
PREFILL_TOKENS = 128


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="rwkv6-7b:Q5_1", help="<preset>:<format>, presets in tools/synthetic_model.py")
    ap.add_argument("--prefill-steps", type=int, default=8)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0, help="CPU seconds the bounded reference sample may take per leg")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
def workload_file(spec, rank=0, world=1, barrier=None):
    import synthetic_model as sm
    preset, fmt = spec.split(":")
    os.makedirs(CACHE_DIR, exist_ok=True)
    path = os.path.join(CACHE_DIR, f"{preset}-{fmt}-seed1.bin")
    if rank == 0 and not os.path.isfile(path):
        tmp = path + ".tmp"
        sm.write_direct(tmp, preset, fmt, seed=1)
        os.replace(tmp, path)
    if barrier:
        barrier()
    return path, sm.PRESETS[preset]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed regions (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self, windows):
        rows = [r for ts, r in self.rows if any(a <= ts <= b for a, b in windows)] or [r for _, r in self.rows]
        sm, reasons, mx = [], set(), None
        for r in rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def host_buffers(n_state, n_logits):
    """Pinned host buffers for the e2e leg (what a serving process would hold); falls back to pageable numpy."""
    try:
        import torch
        if torch.cuda.is_available():
            s = torch.zeros(n_state, dtype=torch.float32).pin_memory()
            l = torch.zeros(n_logits, dtype=torch.float32).pin_memory()
            return s, l, s.data_ptr(), l.data_ptr(), "pinned"
    except Exception:
        pass
    s, l = np.zeros(n_state, np.float32), np.zeros(n_logits, np.float32)
    return s, l, s.ctypes.data, l.ctypes.data, "pageable"


def prefill_matmul_flops(preset, n_tokens):
    """FLOPs of the weight contractions of one chunk (2 * weights * tokens), per-layer matrices only; None for shapes not modelled."""
    if preset.get("arch", (0, 0))[0] != 6:
        return None
    C, F, Lr, mix, dec = preset["C"], preset["F"], preset["L"], preset["mix"], preset["decay"]
    per_layer = 5 * C * C + C * 5 * mix + 5 * mix * C + 2 * C * dec + 2 * C * F + C * C
    return 2.0 * per_layer * Lr * n_tokens


def measured_tensor_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"
    except Exception:
        return 1500.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(preset):
    """DRAM bytes of the dominant kernel's largest per-layer launch from the committed ncu capture (tools/ncu_summary.py --traffic),
    next to that launch's algorithmic bytes (ffn key + receptance batch: (F + C) rows of C/32 Q5_1 blocks)."""
    try:
        import glob
        f = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ncu_traffic.json")))[-1]
        t = json.load(open(f))
        algo = (preset["F"] + preset["C"]) * (preset["C"] // 32) * 24
        return t["dram_bytes"], {"source": os.path.relpath(f, ROOT), "kernel": t["kernel"], "launch": "ffn key + receptance (one launch, 296 CTAs)",
                                 "algorithmic_bytes": algo, "dram_over_algorithmic": t["dram_bytes"] / algo}
    except Exception:
        return None, None


def measured_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def cpu_reference_leg(path, preset, n_decode, budget_s, want_prefill=True):
    """Times the unmodified reference (oracle/_ref) on the host cores: bounded sample of the same workload."""
    import ref_lib
    import synthetic_model as sm
    ref = ref_lib.load_reference_library()
    ref.rwkv_set_print_errors(None, False)
    cores = available_cpus()
    load_s = 0.0
    if "RWKV_REF_THREADS" in os.environ:
        candidates = [int(os.environ["RWKV_REF_THREADS"])]
    else:   # ggml's spin barriers collapse when oversubscribed: probe a few counts, keep the fastest
        candidates = sorted({max(1, min(cores, c)) for c in (8, 16, 32, 64, cores)})
    best = None
    for th in candidates:
        t0 = time.time()
        c = ref.rwkv_init_from_file(path.encode(), th, 0)
        if not c:
            raise RuntimeError("reference failed to load " + path)
        load_s = time.time() - t0
        n_state, n_vocab = ref.rwkv_get_state_len(c), ref.rwkv_get_logits_len(c)
        st, lg = np.zeros(n_state, np.float32), np.zeros(n_vocab, np.float32)
        ref.rwkv_init_state(c, st.ctypes.data_as(PF))
        ref.rwkv_eval(c, 1, st.ctypes.data_as(PF), st.ctypes.data_as(PF), lg.ctypes.data_as(PF))
        t0 = time.perf_counter()
        for t in (2, 3):
            ref.rwkv_eval(c, t, st.ctypes.data_as(PF), st.ctypes.data_as(PF), lg.ctypes.data_as(PF))
        dt = (time.perf_counter() - t0) / 2
        log("reference probe: %d threads -> %.1f ms/token" % (th, dt * 1e3))
        ref.rwkv_free(c)
        if best is None or dt < best[1]:
            best = (th, dt)
        if dt > 4 * best[1]:
            break
    threads = best[0]
    ctx = ref.rwkv_init_from_file(path.encode(), threads, 0)
    n_state, n_vocab = ref.rwkv_get_state_len(ctx), ref.rwkv_get_logits_len(ctx)
    state, logits = np.zeros(n_state, np.float32), np.zeros(n_vocab, np.float32)
    ref.rwkv_init_state(ctx, state.ctypes.data_as(PF))
    toks = sm.synthetic_tokens(4096, n_vocab)
    sp, lp = state.ctypes.data_as(PF), logits.ctypes.data_as(PF)
    # warm-up (2 tokens: first call builds the scheduler) and a probe to size the sample
    t0 = time.time()
    for t in toks[:2]:
        ref.rwkv_eval(ctx, t, sp, sp, lp)
    probe = (time.time() - t0) / 2
    n = int(max(4, min(n_decode, budget_s / max(probe, 1e-6))))
    times = []
    for t in toks[2:2 + n]:
        t0 = time.perf_counter()
        ref.rwkv_eval(ctx, t, sp, sp, lp)
        times.append(time.perf_counter() - t0)
    med = float(np.median(times))
    out = {"decode_tokens_per_s": 1.0 / med, "decode_ms_per_token": med * 1e3, "decode_sample_tokens": n, "threads": threads, "cores": cores,
           "load_s": load_s, "library": os.path.basename(ref_lib.reference_library_path())}
    if want_prefill and med * PREFILL_TOKENS < 6 * budget_s:
        arr = (ctypes.c_uint32 * PREFILL_TOKENS)(*toks[:PREFILL_TOKENS])
        ref.rwkv_eval_sequence_in_chunks(ctx, arr, PREFILL_TOKENS, PREFILL_TOKENS, None, sp, lp)   # builds + caches the sequence graph
        t0 = time.perf_counter()
        ref.rwkv_eval_sequence_in_chunks(ctx, arr, PREFILL_TOKENS, PREFILL_TOKENS, None, sp, lp)
        dt = time.perf_counter() - t0
        out["prefill_tokens_per_s"] = PREFILL_TOKENS / dt
        out["prefill_sample"] = "1 chunk of 128 tokens after one warm-up chunk"
    ref.rwkv_free(ctx)
    return out


def available_cpus():
    """CPUs this process may actually run on: affinity mask, capped by the cgroup quota (a container often sees every
    core of the host in os.cpu_count() while being limited to a few)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period))))
    except (OSError, ValueError):
        pass
    return n


def log(msg):
    print("[bench %6.1fs] %s" % (time.time() - _T0, msg), file=sys.stderr, flush=True)


_T0 = time.time()


def cpu_model_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    path, preset = workload_file(args.workload)
    r = cpu_reference_leg(path, preset, n_decode=args.steps, budget_s=max(args.cpu_budget_s, 10.0))
    line = {
        "impl": "reference", "metric": "decode_tokens_per_sec", "value": r["decode_tokens_per_s"], "unit": "tokens/s", "n_gpus": args.gpus,
        "steps": r["decode_sample_tokens"], "warmup": 2, "ms_per_step": r["decode_ms_per_token"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "q5_1 weights x q8_1 activations (int8 dot, fp32 accumulate)" if "Q" in args.workload else "as file",
        "data": "synthetic", "config": {"workload": args.workload + " single-token rwkv_eval, reference CPU path (oracle/_ref)", "cpu": cpu_model_name()},
        "cpu_baseline": {"value": r["decode_tokens_per_s"], "unit": "tokens/s", "cores": r["threads"], "kind": "reference",
                         "sample": f"{r['decode_sample_tokens']} rwkv_eval calls (median), {r['library']}, {r['threads']} threads of {r['cores']} logical CPUs"},
        "e2e": {"value": r["decode_tokens_per_s"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "prefill": {"tokens_per_s": r.get("prefill_tokens_per_s"), "chunk": PREFILL_TOKENS},
        "gpu_launches": 0,
    }
    emit(line)


def run_ours(args, rank, world, dist):
    import __graft_entry__
    import synthetic_model as sm
    pkg = __graft_entry__.load_package()
    lib = pkg.load_rwkv_shared_library()
    L = lib.library
    barrier = (lambda: dist.barrier()) if dist else None
    path, preset = workload_file(args.workload, rank, world, barrier)
    log("workload file ready: %s" % path)
    local = int(os.environ.get("LOCAL_RANK", rank))
    t0 = time.time()
    ctx = lib.rwkv_b200_init_from_file_ex(path, local, 0, -1)   # N > 1: one full replica per GPU (multi-GPU)
    load_s = time.time() - t0
    log("loaded %s in %.1fs" % (path, load_s))
    n_state, n_vocab = lib.rwkv_get_state_len(ctx), lib.rwkv_get_logits_len(ctx)
    K, W = args.steps, args.warmup
    toks = sm.synthetic_tokens(max(K + W + 8, (args.prefill_steps + 3) * PREFILL_TOKENS), n_vocab)
    tok_arr = (ctypes.c_uint32 * len(toks))(*toks)
    sampler = ClockSampler(local)
    sampler.start()
    windows = []

    def sync_all():
        L.rwkv_b200_synchronize(ctx.ptr)
        if dist:
            dist.barrier()

    # ---- decode, state resident in HBM, CUDA events on the library's stream -------------------------------
    L.rwkv_b200_state_load(ctx.ptr, None)
    L.rwkv_b200_eval_resident(ctx.ptr, tok_arr, 1, True, None)   # lazy allocations, graph capture happens in warm-up
    sync_all()
    launches0 = L.rwkv_b200_kernel_launch_count()
    w0 = time.time()
    ms = L.rwkv_b200_time_resident(ctx.ptr, tok_arr, 1, K, W, True)
    windows.append((w0, time.time()))
    launches = (L.rwkv_b200_kernel_launch_count() - launches0) * K // (K + W)
    assert ms > 0, "device timing failed"
    ms_max = ms
    if dist:
        import torch
        t = torch.tensor([ms], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
    decode_tps = world * K / (ms_max / 1e3)
    persistent = int(L.rwkv_b200_persistent_state(ctx.ptr)) == 1
    overlap_groups = int(L.rwkv_b200_overlap_groups(ctx.ptr))
    log("decode resident: %.3f ms/token (%s)" % (ms_max / K, "one persistent kernel per token" if persistent else "CUDA graph of per-launch kernels"))

    # ---- decode end to end through rwkv_eval with host buffers ------------------------------------------------
    sbuf, lbuf, sp, lp, kind = host_buffers(n_state, n_vocab)
    lib.rwkv_init_state(ctx, sp)
    for t in toks[:W]:
        lib.rwkv_eval(ctx, t, sp, sp, lp)
    sync_all()
    w0 = time.time()
    t0 = time.perf_counter()
    for t in toks[W:W + K]:
        lib.rwkv_eval(ctx, t, sp, sp, lp)
    e2e_s = time.perf_counter() - t0
    windows.append((w0, time.time()))
    if dist:
        import torch
        t = torch.tensor([e2e_s], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_tps = world * K / e2e_s
    log("decode e2e: %.3f ms/token" % (e2e_s / K * 1e3))

    # ---- prefill: 128-token chunk ----------------------------------------------------------------------------------
    P = args.prefill_steps
    L.rwkv_b200_state_load(ctx.ptr, None)
    w0 = time.time()
    pms = L.rwkv_b200_time_resident(ctx.ptr, tok_arr, PREFILL_TOKENS, P, 2, True)
    windows.append((w0, time.time()))
    arr128 = (ctypes.c_uint32 * PREFILL_TOKENS)(*toks[:PREFILL_TOKENS])
    L.rwkv_eval_sequence_in_chunks(ctx.ptr, arr128, PREFILL_TOKENS, PREFILL_TOKENS, None, ctypes.cast(sp, PF), ctypes.cast(lp, PF))
    t0 = time.perf_counter()
    for _ in range(max(2, P // 2)):
        L.rwkv_eval_sequence_in_chunks(ctx.ptr, arr128, PREFILL_TOKENS, PREFILL_TOKENS, None, ctypes.cast(sp, PF), ctypes.cast(lp, PF))
    pe2e_s = (time.perf_counter() - t0) / max(2, P // 2)
    if dist:
        import torch
        t = torch.tensor([pms, pe2e_s], device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        pms, pe2e_s = float(t[0].item()), float(t[1].item())

    log("prefill: %.3f ms/chunk resident, %.3f ms/chunk e2e" % (pms / P, pe2e_s * 1e3))
    # ---- roofline leg: events around every GEMV launch of one decode pass ------------------------------------
    prof = pkg.shared_library.ProfileResult()
    L.rwkv_b200_state_load(ctx.ptr, None)
    profs = []
    for i in range(5):
        L.rwkv_b200_profile_pass(ctx.ptr, ctypes.cast(ctypes.byref(tok_arr, 4 * i), PU), 1, True, ctypes.byref(prof))
        profs.append((prof.gemv_ms, prof.gemv_bytes, prof.pass_ms, prof.gemv_launches, prof.total_launches, prof.top_ms, prof.top_bytes))
    profs = profs[1:]
    gemv_ms = float(np.median([p[0] for p in profs])); gemv_bytes = profs[0][1]; pass_ms = float(np.median([p[2] for p in profs]))
    top = min(profs, key=lambda p: p[5])
    peak, peak_src = measured_peaks()
    bytes_tok = int(L.rwkv_b200_bytes_per_token(ctx.ptr, True))
    step_ms = ms_max / K
    sampler.stop()
    clocks = sampler.summary(windows)
    log("roofline leg done")

    line = None
    if rank == 0:
        line = {
            "metric": "decode_tokens_per_sec", "value": decode_tps, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "q5_1 weights x q8_1 activations (int8 dp4a, fp32 accumulate); fp16 head" if "Q5_1" in args.workload else args.workload.split(":")[1],
            "data": "synthetic",
            "config": {"workload": f"{args.workload} single-token rwkv_eval with logits ({preset['L']} layers, n_embed {preset['C']}, ffn {preset['F']}, vocab {preset['V']})",
                       "parallelism": "1 GPU" if world == 1 else f"{world} independent replicas, one stream each (single-stream decode does not shard; DESIGN.md)",
                       "l2": "6.1 GB of weights per step >> 126 MB L2, no flush needed", "bytes_per_token": bytes_tok, "load_s": round(load_s, 2),
                       "cuda_graph": not persistent, "persistent_kernel": persistent},
            "clocks": clocks,
            "e2e": {"value": e2e_tps, "unit": "tokens/s", "h2d_bytes_per_step": n_state * 4 + 4, "d2h_bytes_per_step": n_state * 4 + n_vocab * 4,
                    "host_buffers": kind, "ms_per_step": e2e_s / K * 1e3,
                    "state_copies": ("pipelined against %d layer groups on copy streams" % overlap_groups) if overlap_groups else "one upload before, one download after the pass"},
            "gpu_launches": int(launches),
            "prefill": {"tokens_per_s": world * P * PREFILL_TOKENS / (pms / 1e3), "ms_per_chunk": pms / P, "chunk": PREFILL_TOKENS, "steps": P,
                        "e2e_tokens_per_s": world * PREFILL_TOKENS / pe2e_s,
                        "kernel": "gemm_tc_kernel (tcgen05, TMEM accumulators + TMEM A operand) for every layer matrix; wkv6 / lerp / LN on CUDA cores",
                        "tensor": (lambda fl, pk: None if fl is None else {"bound": "tensor", "achieved": fl / (pms / P * 1e-3) / 1e12, "peak": pk[0], "unit": "TFLOP/s",
                                                                            "frac": fl / (pms / P * 1e-3) / 1e12 / pk[0], "peak_source": pk[1],
                                                                            "note": "whole chunk (GEMMs + recurrence + glue) against the dense bf16 peak"})(
                            prefill_matmul_flops(preset, PREFILL_TOKENS), measured_tensor_peak())},
            "roofline": {"bound": "hbm", "kernel": "gemv_kernel (fused dequantize-GEMV, all launches of one decode step)",
                         "achieved": gemv_bytes / (gemv_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": gemv_bytes / (gemv_ms * 1e-3) / 1e9 / peak,
                         "peak_source": peak_src, "traffic": ncu_traffic(preset)[0] if "rwkv6-7b:Q5_1" == args.workload else None,
                         "traffic_launch": ncu_traffic(preset)[1] if "rwkv6-7b:Q5_1" == args.workload else None,
                         "bytes_per_step": gemv_bytes, "ms_per_step": gemv_ms,
                         "launches_per_step": int(profs[0][3]), "share_of_step": gemv_ms / pass_ms,
                         "largest_launch": {"bytes": top[6], "ms": top[5], "gbs": top[6] / (top[5] * 1e-3) / 1e9},
                         "whole_step": {"bytes": bytes_tok, "ms": step_ms, "gbs": bytes_tok / (step_ms * 1e-3) / 1e9, "frac": bytes_tok / (step_ms * 1e-3) / 1e9 / peak}},
        }
    if line is not None and persistent:
        # the whole token is ONE kernel (plus the embedding gather): that kernel is the dominant kernel, its algorithmic bytes are the
        # byte model of a token and its duration the CUDA-event time of a step; the per-launch GEMV figures (measured
        # with the per-launch path, which the profiling leg always uses) stay for comparison.
        r = line["roofline"]
        per_launch = {k: r[k] for k in ("kernel", "achieved", "frac", "bytes_per_step", "ms_per_step", "launches_per_step", "share_of_step", "largest_launch")}
        r.update({"kernel": "decode_persistent_kernel (every layer's fused dequantize-GEMVs, LayerNorm/mix, lerp and WKV of one token in one launch)",
                  "achieved": r["whole_step"]["gbs"], "frac": r["whole_step"]["frac"], "bytes_per_step": bytes_tok, "ms_per_step": step_ms,
                  "launches_per_step": 1, "share_of_step": 1.0, "per_launch_path": per_launch})
        r.pop("largest_launch", None)
    lib.rwkv_free(ctx)
    if rank == 0:
        if world == 1 and not args.skip_cpu_baseline:
            # the reference runs in its own process under a hard timeout: it can never hang or skew the GPU numbers
            log("cpu baseline (reference arm in a subprocess)")
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", args.workload, "--steps", "32",
                                    "--cpu-budget-s", str(args.cpu_budget_s)], capture_output=True, text=True, timeout=max(240.0, 14 * args.cpu_budget_s))
                ref_line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
                line["cpu_baseline"] = dict(ref_line["cpu_baseline"], prefill_tokens_per_s=ref_line["prefill"]["tokens_per_s"], cpu=ref_line["config"]["cpu"])
            except Exception as e:   # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}
        emit(line)


def run_pipeline(args, rank, world, dist):
    """N > 1: layer pipeline (rwkv.cpp_b200/pipeline.py). Rank g holds layers [g*L/N, (g+1)*L/N) and their slice of
    the recurrent state; `world` sequences are in flight, one per stage; x f32[C x T] crosses each stage boundary by NCCL
    send/recv on the stream the stage kernels run on. One step = one token of every in-flight sequence (`world` tokens)."""
    import torch
    import __graft_entry__
    import synthetic_model as sm
    pkg = __graft_entry__.load_package()
    pipeline = pkg.pipeline
    lib = pkg.load_rwkv_shared_library()
    L = lib.library
    path, preset = workload_file(args.workload, rank, world, lambda: dist.barrier())
    local = int(os.environ.get("LOCAL_RANK", rank))
    begin, end = pipeline.stage_layers(preset["L"], world, rank)
    t0 = time.time()
    ctxs = [lib.rwkv_b200_init_from_file_ex(path, local, begin, end)]
    for _ in range(world - 1):
        ctxs.append(lib.rwkv_clone_context(ctxs[0], 1))          # one context = one in-flight sequence's state slice
    load_s = time.time() - t0
    log("rank %d: layers [%d, %d) loaded in %.1fs" % (rank, begin, end, load_s))
    n_vocab = lib.rwkv_get_logits_len(ctxs[0])
    for c in ctxs:
        L.rwkv_b200_state_load(c.ptr, None)
        L.rwkv_b200_synchronize(c.ptr)
    K, W, P = args.steps, args.warmup, args.prefill_steps
    first, last = rank == 0, rank == world - 1
    # a dedicated (non-default) stream: torch's NCCL send / recv and the stage kernels are all ordered on it. (The legacy default
    # stream has handle 0, which rwkv_b200_stage_eval reads as "use the context's own stream".)
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    sp = ctypes.c_void_p(stream.cuda_stream)
    assert stream.cuda_stream != 0
    transport = pipeline.Transport(dist, rank, world)
    sampler = ClockSampler(local)
    sampler.start()
    windows = []
    logits_host = torch.zeros(n_vocab, dtype=torch.float32).pin_memory()
    bytes_tok = int(L.rwkv_b200_bytes_per_token(ctxs[0].ptr, True))   # this stage's share of the byte model

    def leg(T, steps, warm, read_logits):
        """`steps` timed steps of `world` work items each (one per in-flight sequence), after `warm` untimed ones."""
        toks = sm.synthetic_tokens((steps + warm) * world * T + T, n_vocab)
        arr = (ctypes.c_uint32 * len(toks))(*toks)
        n_hidden = L.rwkv_b200_stage_hidden_len(ctxs[0].ptr, T)
        recv = torch.empty(n_hidden, dtype=torch.float32, device=f"cuda:{local}")
        send = torch.empty(n_hidden, dtype=torch.float32, device=f"cuda:{local}")
        state = {"u": 0}

        def stage(seq, step, hin, hout):
            u = state["u"]; state["u"] += 1
            ok = L.rwkv_b200_stage_eval(ctxs[seq].ptr, ctypes.cast(ctypes.byref(arr, 4 * u * T), PU), T,
                                        ctypes.c_void_p(hin.data_ptr()) if hin is not None else None,
                                        ctypes.c_void_p(hout.data_ptr()) if hout is not None else None, True, sp)
            assert ok, "stage_eval failed"
            if read_logits and last:
                assert L.rwkv_b200_stage_logits(ctxs[seq].ptr, ctypes.cast(logits_host.data_ptr(), PF), sp)

        pipeline.run_ticks(transport, pipeline.schedule(world, world, warm * world, rank), stage, recv, send)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = L.rwkv_b200_kernel_launch_count()
        w0 = time.time(); t0 = time.perf_counter()
        e0.record(stream)
        pipeline.run_ticks(transport, pipeline.schedule(world, world, steps * world, rank), stage, recv, send)
        e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dist.barrier()
        windows.append((w0, time.time()))
        t = torch.tensor([e0.elapsed_time(e1), wall * 1e3, float(L.rwkv_b200_kernel_launch_count() - launches0)], device=f"cuda:{local}")
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm_ = t.clone(); dist.all_reduce(sm_, op=dist.ReduceOp.SUM)
        return float(mx[0]), float(mx[1]), float(sm_[2])

    dev_ms, _, launches = leg(1, K, max(W, 3), False)                      # decode, logits stay on the last stage's GPU
    log("pipeline decode: %.3f ms per step of %d tokens" % (dev_ms / K, world))
    _, e2e_ms, _ = leg(1, K, max(W, 3), True)                              # + token H2D on stage 0, logits D2H on the last stage, per token
    pdev_ms, _, _ = leg(PREFILL_TOKENS, P, 2, False)
    _, pe2e_ms, _ = leg(PREFILL_TOKENS, P, 1, True)
    sampler.stop()
    clocks = sampler.summary(windows)
    stage_bytes = torch.tensor([float(bytes_tok)], device=f"cuda:{local}")
    dist.all_reduce(stage_bytes, op=dist.ReduceOp.MAX)
    for c in ctxs[1:]:
        lib.rwkv_free(c)
    lib.rwkv_free(ctxs[0])
    if rank == 0:
        peak, peak_src = measured_peaks()
        tick_ms = dev_ms / (K * world)                                      # one token leaves the pipeline per tick
        gbs = float(stage_bytes.item()) / (tick_ms * 1e-3) / 1e9
        line = {
            "metric": "decode_tokens_per_sec", "value": world * K / (dev_ms / 1e3), "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": max(W, 3),
            "ms_per_step": dev_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "q5_1 weights x q8_1 activations (int8 dp4a, fp32 accumulate); fp16 head" if "Q5_1" in args.workload else args.workload.split(":")[1],
            "data": "synthetic",
            "config": {"workload": f"{args.workload} single-token eval with logits ({preset['L']} layers, n_embed {preset['C']}, ffn {preset['F']}, vocab {preset['V']})",
                       "parallelism": f"pp{world}: {preset['L'] // world} layers + their state slice per GPU, {world} sequences in flight (one per stage), "
                                      f"x f32[{preset['C']}] per token over NCCL send/recv; one step = {world} tokens",
                       "l2": "each stage streams its share of 6.1 GB of weights per token >> 126 MB L2, no flush needed", "load_s": round(load_s, 2), "cuda_graph": True},
            "clocks": clocks,
            "e2e": {"value": world * K / (e2e_ms / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": 4 * world, "d2h_bytes_per_step": 4 * n_vocab * world,
                    "host_buffers": "pinned", "ms_per_step": e2e_ms / K,
                    "note": "token ids enter on stage 0, logits leave from the last stage; the recurrent state stays resident on its stage"},
            "gpu_launches": int(launches),
            "prefill": {"tokens_per_s": world * P * PREFILL_TOKENS / (pdev_ms / 1e3), "ms_per_chunk": pdev_ms / (P * world), "chunk": PREFILL_TOKENS, "steps": P,
                        "e2e_tokens_per_s": world * P * PREFILL_TOKENS / (pe2e_ms / 1e3)},
            "roofline": {"bound": "hbm", "kernel": "whole pipeline tick of the slowest stage (fused dequant-GEMV launches + glue)",
                         "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "peak_source": peak_src, "traffic": None,
                         "bytes_per_step": float(stage_bytes.item()), "ms_per_step": tick_ms,
                         "note": "per GPU: the largest stage's bytes per token / time per pipeline tick; per-launch figures are in the N=1 line"},
            "cpu_baseline": {"value": None, "unit": "tokens/s", "cores": 0, "kind": "reference", "sample": "reported by the N=1 run and by --impl reference"},
        }
        emit(line)


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's real stdout; everything libraries print (NCCL's version banner ...) goes to stderr."""
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    dist = None
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        dist_mod.init_process_group("nccl")
        dist = dist_mod
    if world > 1:
        run_pipeline(args, rank, world, dist)
    else:
        run_ours(args, rank, world, dist)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
