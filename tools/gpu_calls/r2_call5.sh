#!/bin/bash
# Round 2, GPU call 5: the rewritten prefill GEMM (fp16 operand blocks expanded at load, bulk-copy ring, SS tcgen05.mma): unit tests,
# parity, prefill bench + timeline; full-shape tests against the regenerated (damped-preset) goldens.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tensor-core unit tests"
timeout 600 $PY -m pytest tests/test_gpu_gemv.py -q -m gpu --timeout 300 -k "tensor_core" --maxfail 30 -rfE > gpurun_out/r2_c5_tc_unit.log 2>&1; echo "tc unit rc=$?"; tail -n 2 gpurun_out/r2_c5_tc_unit.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c5_tc_unit.log | head -20; grep -E "^E +Assertion" gpurun_out/r2_c5_tc_unit.log | cut -c1-200 | head -12
echo "== 2. core tests"
for f in gemv parity pipeline batch; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c5_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c5_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c5_$f.log | head -14
done
echo "== 3. prefill"
pf() { name=$1; shift; env "$@" timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c5_pf_$name.json 2> gpurun_out/r2_c5_pf_$name.log; echo "$name rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c5_pf_$name.log | tail -1)"; }
pf default RWKV_B200_X=0
pf nosplit RWKV_B200_TC_SPLITK=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c5.csv > gpurun_out/r2_trace_prefill_c5.log 2>&1; tail -n 45 gpurun_out/r2_trace_prefill_c5.log
echo "== 4. full shape"
timeout 1200 $PY -m pytest tests/test_gpu_full_shape.py -q -m gpu --timeout 600 -rfE -s > gpurun_out/r2_c5_full_shape.log 2>&1; echo "full shape rc=$?"; tail -n 3 gpurun_out/r2_c5_full_shape.log; grep -E "max\|ours|tensor-core vs|^(FAILED|ERROR)" gpurun_out/r2_c5_full_shape.log | cut -c1-220 | head -30
ls gpurun_out | grep c5 | head
