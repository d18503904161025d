#!/bin/bash
# Round 2, GPU call 12: tail job removed (ln_mix kernel emits the staged columns instead), K-splits of the prefill GEMM reduced through
# distributed shared memory of a thread-block cluster: tests, decode A/B, prefill bench + timeline.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
for f in gemv parity; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c12_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c12_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c12_$f.log | head -14
done
echo "== 2. A/B decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 64 > gpurun_out/r2_c12_ab_$name.json 2> gpurun_out/r2_c12_ab_$name.log; echo "$name rc=$? $(cut -c1-330 gpurun_out/r2_c12_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab noxq RWKV_B200_NO_XQ=1
echo "== 3. prefill"
pf() { name=$1; shift; env "$@" timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c12_pf_$name.json 2> gpurun_out/r2_c12_pf_$name.log; echo "$name rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c12_pf_$name.log | tail -1)"; }
pf default RWKV_B200_X=0
pf nosplit RWKV_B200_TC_SPLITK=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c12.csv > gpurun_out/r2_trace_prefill_c12.log 2>&1; tail -n 42 gpurun_out/r2_trace_prefill_c12.log
du -sh gpurun_out
