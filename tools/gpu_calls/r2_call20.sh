#!/bin/bash
# Round 2, GPU call 20: batched staging loads in the cluster reduction, wkv6 full chunks as one software-pipelined block.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
for f in parity batch; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c20_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c20_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c20_$f.log | head -14
done
timeout 300 $PY -m pytest tests/test_gpu_gemv.py -q -m gpu --timeout 300 -k "gemm or tc or tensor" > gpurun_out/r2_c20_gemm.log 2>&1; echo "gemm unit rc=$?"; tail -n 1 gpurun_out/r2_c20_gemm.log
echo "== 2. prefill"
timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c20_pf_default.json 2> gpurun_out/r2_c20_pf_default.log; echo "rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c20_pf_default.log | tail -1)"; grep -E "decode resident" gpurun_out/r2_c20_pf_default.log
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c20.csv > gpurun_out/r2_trace_prefill_c20.log 2>&1; tail -n 32 gpurun_out/r2_trace_prefill_c20.log
du -sh gpurun_out
