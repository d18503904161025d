#!/bin/bash
# Round 2, GPU call 15: pageable buffers through a persistent pool of copy lanes; phase marks inside the prefill GEMM.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
nproc; grep -c processor /proc/cpuinfo; cat /sys/fs/cgroup/cpu.max 2>/dev/null
echo "== 1. tests"
for f in overlap parity; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c15_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c15_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c15_$f.log | head -14
done
echo "== 2. bench line with e2e pinned / pageable"
timeout 400 $PY bench.py --steps 64 --skip-cpu-baseline > gpurun_out/r2_c15_bench.json 2> gpurun_out/r2_c15_bench.log; echo "rc=$?"; grep -E "decode|e2e|prefill" gpurun_out/r2_c15_bench.log | tail -8
echo "== 3. prefill timeline with GEMM phase marks"
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c15.csv > gpurun_out/r2_trace_prefill_c15.log 2>&1; tail -n 34 gpurun_out/r2_trace_prefill_c15.log
du -sh gpurun_out
