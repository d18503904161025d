#!/bin/bash
# Round 2, GPU call 16: per-format instantiations of the decode GEMV (instruction footprint), compact rolled GEMM epilogue,
# residual / gate inputs of the split-K reduction prefetched before the cluster barrier.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
for f in gemv parity batch; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c16_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c16_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c16_$f.log | head -14
done
echo "== 2. A/B decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 64 > gpurun_out/r2_c16_ab_$name.json 2> gpurun_out/r2_c16_ab_$name.log; echo "$name rc=$? $(cut -c1-330 gpurun_out/r2_c16_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab generic_kernel RWKV_B200_GEMV_PER_TYPE=0
ab default2 RWKV_B200_X=1
echo "== 3. prefill"
timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c16_pf_default.json 2> gpurun_out/r2_c16_pf_default.log; echo "rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c16_pf_default.log | tail -1)"
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c16.csv > gpurun_out/r2_trace_prefill_c16.log 2>&1; tail -n 32 gpurun_out/r2_trace_prefill_c16.log
du -sh gpurun_out
