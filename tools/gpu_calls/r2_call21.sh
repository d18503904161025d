#!/bin/bash
# Round 2, GPU call 21: ln_mix_kernel writes the fp16 GEMM operands of its consumers (no convert_f16 launch in front of those GEMMs).
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
timeout 900 $PY -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c21_parity.log 2>&1; echo "parity rc=$?"; tail -n 2 gpurun_out/r2_c21_parity.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c21_parity.log | head -14
timeout 900 $PY -m pytest tests/test_gpu_full_shape.py -q -m gpu --timeout 600 -k "chunked_prefill" -rfE -s > gpurun_out/r2_c21_full_shape.log 2>&1; echo "full shape rc=$?"; grep -E "passed|failed|max\|ours" gpurun_out/r2_c21_full_shape.log | tail -8
echo "== 2. prefill A/B"
pf() { name=$1; shift; env "$@" timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c21_pf_$name.json 2> gpurun_out/r2_c21_pf_$name.log; echo "$name rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c21_pf_$name.log | tail -1) $(grep -o 'decode resident: [0-9.]* ms' gpurun_out/r2_c21_pf_$name.log)"; }
pf default RWKV_B200_X=0
pf noln16 RWKV_B200_NO_LN16=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c21.csv > gpurun_out/r2_trace_prefill_c21.log 2>&1; grep -A8 "critical-path" gpurun_out/r2_trace_prefill_c21.log
du -sh gpurun_out
