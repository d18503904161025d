#!/bin/bash
# Round 2, GPU call 10 (= call 9 re-run after the session that issued it was lost): LayerNorm as the tail job of the x-writing GEMV +
# staged columns from lerp / WKV / tail: tests, A/B, timeline; memcheck of the in-process pipeline (event-based completion).
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. core tests"
for f in gemv parity batch pipeline overlap; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c10_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c10_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c10_$f.log | head -14
done
echo "== 2. A/B decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 64 > gpurun_out/r2_c10_ab_$name.json 2> gpurun_out/r2_c10_ab_$name.log; echo "$name rc=$? $(cut -c1-330 gpurun_out/r2_c10_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab notail RWKV_B200_NO_LN_TAIL=1
ab noxq RWKV_B200_NO_XQ=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --out gpurun_out/r2_trace_decode_c10.csv > gpurun_out/r2_trace_decode_c10.log 2>&1; tail -n 40 gpurun_out/r2_trace_decode_c10.log
echo "== 3. prefill"
timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c10_pf_default.json 2> gpurun_out/r2_c10_pf_default.log; echo "rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c10_pf_default.log | tail -1)"
echo "== 4. sanitizer"
timeout 240 compute-sanitizer --tool memcheck --print-limit 20 --error-exitcode 9 $PY -m pytest tests/test_gpu_pipeline.py -q -m gpu -x > gpurun_out/r2_c10_memcheck_pipeline.log 2>&1; echo "memcheck pipeline rc=$?"; tail -n 3 gpurun_out/r2_c10_memcheck_pipeline.log
du -sh gpurun_out
