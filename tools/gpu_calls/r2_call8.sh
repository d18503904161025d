#!/bin/bash
# Round 2, GPU call 8: validate the slimmed kernels (PRO_LN_MIX / fused decay removed, tiny GEMM batches on the GEMV path), A/B full tiles,
# compute-sanitizer on the fixture suite, the other BASELINE configurations.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. core tests"
for f in gemv parity batch pipeline overlap; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c8_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c8_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c8_$f.log | head -14
done
echo "== 2. A/B decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 64 > gpurun_out/r2_c8_ab_$name.json 2> gpurun_out/r2_c8_ab_$name.log; echo "$name rc=$? $(cut -c1-330 gpurun_out/r2_c8_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab tiles_wr RWKV_B200_FULL_TILES=0
ab default2 RWKV_B200_X=1
echo "== 3. A/B prefill"
pf() { name=$1; shift; env "$@" timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c8_pf_$name.json 2> gpurun_out/r2_c8_pf_$name.log; echo "$name rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c8_pf_$name.log | tail -1)"; }
pf default RWKV_B200_TC_SPLITK=1
pf split RWKV_B200_X=0
pf alltc RWKV_B200_TC_MIN_WEIGHTS=0 RWKV_B200_TC_SPLITK=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c8.csv > gpurun_out/r2_trace_prefill_c8.log 2>&1; grep -A8 "critical-path" gpurun_out/r2_trace_prefill_c8.log
echo "== 4. compute-sanitizer"
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 $PY -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "fixture_logits or long_prompt or two_clones or large_activ" > gpurun_out/r2_c8_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -n 4 gpurun_out/r2_c8_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 --error-exitcode 9 $PY -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "long_prompt and (6v0 or 7v0 or 4v0) and (Q5_1 or FP16)" > gpurun_out/r2_c8_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -n 4 gpurun_out/r2_c8_racecheck.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 $PY -m pytest tests/test_gpu_batch.py tests/test_sampling.py tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/r2_c8_memcheck_batch.log 2>&1; echo "memcheck batch/pipe rc=$?"; tail -n 3 gpurun_out/r2_c8_memcheck_batch.log
echo "== 5. other configs"
timeout 300 $PY bench.py --workload rwkv4-169m:Q5_1 --steps 256 > gpurun_out/r2_c8_bench_169m.json 2> gpurun_out/r2_c8_bench_169m.log; echo "169m rc=$? $(cut -c1-260 gpurun_out/r2_c8_bench_169m.json)"
timeout 400 $PY bench.py --workload rwkv5-1b5:Q4_0 --mode prefill --steps 16 > gpurun_out/r2_c8_bench_1b5_prefill.json 2> gpurun_out/r2_c8_bench_1b5_prefill.log; echo "1b5 rc=$? $(cut -c1-260 gpurun_out/r2_c8_bench_1b5_prefill.json)"
timeout 400 $PY bench.py --workload rwkv7-2b9:FP16 > gpurun_out/r2_c8_bench_2b9.json 2> gpurun_out/r2_c8_bench_2b9.log; echo "2b9 rc=$? $(cut -c1-260 gpurun_out/r2_c8_bench_2b9.json)"
du -sh gpurun_out
