#!/bin/bash
# Round 2, GPU call 6: validate Q5 device-order qh + full-stage rotating tiles + tiled lerp + carve-out pinning + K-staggered GEMM,
# A/B them on decode and prefill, ncu --set full of every kernel.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. core tests"
for f in gemv parity batch pipeline; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c6_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c6_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c6_$f.log | head -14
done
echo "== 2. A/B decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 48 > gpurun_out/r2_c6_ab_$name.json 2> gpurun_out/r2_c6_ab_$name.log; echo "$name rc=$? $(cut -c1-330 gpurun_out/r2_c6_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab nocarve RWKV_B200_CARVEOUT=0
ab nopdl RWKV_B200_NO_PDL=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --out gpurun_out/r2_trace_decode_c6.csv > gpurun_out/r2_trace_decode_c6.log 2>&1; tail -n 36 gpurun_out/r2_trace_decode_c6.log
echo "== 3. A/B prefill"
pf() { name=$1; shift; env "$@" timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c6_pf_$name.json 2> gpurun_out/r2_c6_pf_$name.log; echo "$name rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c6_pf_$name.log | tail -1)"; }
pf default RWKV_B200_X=0
pf nostagger RWKV_B200_TC_STAGGER=0
pf nocarve RWKV_B200_CARVEOUT=0
pf nosplit RWKV_B200_TC_SPLITK=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c6.csv > gpurun_out/r2_trace_prefill_c6.log 2>&1; tail -n 45 gpurun_out/r2_trace_prefill_c6.log
echo "== 4. ncu every kernel"
timeout 900 ncu --set full --clock-control none --import-source on -c 140 -f -o gpurun_out/r2_ncu_all $PY tools/ncu_targets.py > gpurun_out/r2_c6_ncu_all.log 2>&1; echo "ncu all rc=$?"; tail -n 2 gpurun_out/r2_c6_ncu_all.log; ls -la gpurun_out/r2_ncu_all.ncu-rep
ls gpurun_out | grep c6 | head -40
