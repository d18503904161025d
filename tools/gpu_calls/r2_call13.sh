#!/bin/bash
# Round 2, GPU call 13: wkv6 with dedicated normalising warps, lerp unroll, pageable caller buffers through pinned bounce buffers,
# ln_mix emission reverted: tests, decode A/B, e2e (pinned / pageable), prefill bench + timeline.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
for f in parity overlap batch pipeline; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c13_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c13_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c13_$f.log | head -14
done
echo "== 2. decode: bench line with e2e pinned / pageable"
timeout 400 $PY bench.py --steps 64 --skip-cpu-baseline > gpurun_out/r2_c13_bench.json 2> gpurun_out/r2_c13_bench.log; echo "rc=$? $(cut -c1-200 gpurun_out/r2_c13_bench.json)"; grep -E "decode|e2e|prefill" gpurun_out/r2_c13_bench.log | tail -8
echo "== 3. prefill timeline"
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c13.csv > gpurun_out/r2_trace_prefill_c13.log 2>&1; grep -A8 "critical-path" gpurun_out/r2_trace_prefill_c13.log; grep "wkv6 " gpurun_out/r2_trace_prefill_c13.log | tail -2
du -sh gpurun_out
