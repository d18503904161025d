#!/bin/bash
# Round 2, GPU call 1: where round 1 left off, measured. Full -m gpu suite (with the new reference-pinned full-shape and
# >= 32-token tests), default bench + the per-block activation staging variant, the other BASELINE configs, the reference arm,
# compute-sanitizer on the fixture tests, ncu --set full of EVERY kernel + the launch list of one bench run.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
{ nvidia-smi -L; nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv; echo "nproc $(nproc)"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | egrep "Model name|^CPU\(s\)|Thread|Core|Socket"; free -g | head -2; } > gpurun_out/r2_gpu_box.txt 2>&1

echo "== 0. smoke + A/B of the decode variants (quick bench: resident decode + timeline)"
timeout 300 $PY __graft_entry__.py smoke 2>&1 | tail -n 2
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 48 > gpurun_out/r2_c1_ab_$name.json 2> gpurun_out/r2_c1_ab_$name.log; echo "$name rc=$? $(cut -c1-230 gpurun_out/r2_c1_ab_$name.json)"; }
ab base RWKV_B200_NO_FUSE_LN=1
ab fuse RWKV_B200_X=0
ab base_v2 RWKV_B200_NO_FUSE_LN=1 RWKV_B200_STAGE_V2=1
ab fuse_v2 RWKV_B200_STAGE_V2=1
ab fuse_v2_pf4 RWKV_B200_STAGE_V2=1 RWKV_B200_L2_PREFETCH=4
ab fuse_v2_pf16 RWKV_B200_STAGE_V2=1 RWKV_B200_L2_PREFETCH=16
ab fuse_pf8 RWKV_B200_L2_PREFETCH=8

echo "== 1. suite"; timeout 1500 $PY -m pytest tests -q -m gpu --timeout 600 -rA > gpurun_out/r2_c1_suite.log 2>&1; echo "suite rc=$?"; tail -n 3 gpurun_out/r2_c1_suite.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c1_suite.log | head -30
RWKV_B200_NO_FUSE_LN=1 timeout 900 $PY -m pytest tests/test_gpu_parity.py tests/test_gpu_full_shape.py -q -m gpu --timeout 600 > gpurun_out/r2_c1_suite_nofuse.log 2>&1; echo "no-fuse parity rc=$?"; tail -n 2 gpurun_out/r2_c1_suite_nofuse.log
grep -E "max\|ours|tensor-core vs" gpurun_out/r2_c1_suite.log | head -40

echo "== 2. bench (default, 7B Q5_1)"; timeout 600 $PY bench.py > gpurun_out/r2_c1_bench_7b.json 2> gpurun_out/r2_c1_bench_7b.log; echo "bench rc=$?"; tail -n 4 gpurun_out/r2_c1_bench_7b.log
RWKV_B200_STAGE_V2=1 timeout 400 $PY -m pytest tests/test_gpu_gemv.py tests/test_gpu_parity.py -q -x -m gpu > gpurun_out/r2_c1_stagev2_tests.log 2>&1; echo "stage_v2 tests rc=$?"; tail -n 2 gpurun_out/r2_c1_stagev2_tests.log

echo "== 3. other configs"
timeout 300 $PY bench.py --workload rwkv4-169m:Q5_1 --steps 256 > gpurun_out/r2_c1_bench_169m.json 2> gpurun_out/r2_c1_bench_169m.log; echo "169m rc=$?"
timeout 400 $PY bench.py --workload rwkv5-1b5:Q4_0 --mode prefill --steps 16 > gpurun_out/r2_c1_bench_1b5_prefill.json 2> gpurun_out/r2_c1_bench_1b5_prefill.log; echo "1b5 rc=$?"
timeout 400 $PY bench.py --workload rwkv7-2b9:FP16 > gpurun_out/r2_c1_bench_2b9.json 2> gpurun_out/r2_c1_bench_2b9.log; echo "2b9 rc=$?"
timeout 400 $PY bench.py --mode prefill --steps 16 --skip-cpu-baseline > gpurun_out/r2_c1_bench_7b_prefill.json 2> gpurun_out/r2_c1_bench_7b_prefill.log; echo "7b prefill rc=$?"

echo "== 4. reference arm"; ( time timeout 400 $PY bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r2_c1_bench_ref.json 2> gpurun_out/r2_c1_bench_ref.log; echo "ref rc=$?"; tail -n 6 gpurun_out/r2_c1_bench_ref.log

echo "== 5. compute-sanitizer"
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 9 $PY -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fixture_logits or long_prompt or chunked_equals or two_clones or large_activ" > gpurun_out/r2_c1_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 4 gpurun_out/r2_c1_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 9 $PY -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fixture_logits and (6v0 or 7v0 or 4v0) and (Q5_1 or FP16)" > gpurun_out/r2_c1_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 gpurun_out/r2_c1_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 9 $PY -m pytest tests/test_gpu_batch.py tests/test_sampling.py -q -m gpu -x > gpurun_out/r2_c1_memcheck_batch.log 2>&1; echo "memcheck batch rc=$?"; tail -n 3 gpurun_out/r2_c1_memcheck_batch.log

echo "== 6. ncu"
timeout 900 ncu --set full --clock-control none -c 160 -f -o gpurun_out/r2_ncu_all $PY tools/ncu_targets.py > gpurun_out/r2_c1_ncu_all.log 2>&1; echo "ncu all rc=$?"; tail -n 2 gpurun_out/r2_c1_ncu_all.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_ncu_launches.csv $PY bench.py --steps 2 --warmup 1 --prefill-steps 1 --skip-cpu-baseline > gpurun_out/r2_c1_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --out gpurun_out/r2_trace_decode_c1.csv > gpurun_out/r2_trace_decode_c1.log 2>&1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c1.csv > gpurun_out/r2_trace_prefill_c1.log 2>&1
ls -la gpurun_out | tail -30
