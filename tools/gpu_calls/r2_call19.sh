#!/bin/bash
# Round 2, GPU call 19: lean templated split-K reduction with smem-staged residual / gate, GEMV instantiations without the LayerNorm code.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
for f in gemv parity; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c19_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c19_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c19_$f.log | head -14
done
echo "== 2. prefill"
timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c19_pf_default.json 2> gpurun_out/r2_c19_pf_default.log; echo "rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c19_pf_default.log | tail -1)"; grep -E "decode resident" gpurun_out/r2_c19_pf_default.log
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c19.csv > gpurun_out/r2_trace_prefill_c19.log 2>&1; tail -n 32 gpurun_out/r2_trace_prefill_c19.log
du -sh gpurun_out
echo "== 3. decode"
timeout 200 $PY bench.py --quick --steps 64 > gpurun_out/r2_c19_ab_default.json 2> gpurun_out/r2_c19_ab_default.log; echo "rc=$? $(cut -c1-330 gpurun_out/r2_c19_ab_default.json)"
