#!/bin/bash
# Round 2, GPU call 11: after the fix of call 10's regression (kernel parameter copied to local memory by every thread; tail job's
# parameter loads): parity, A/B of the tail job and the staged-column hand-off, timeline.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
for f in parity gemv; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c11_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c11_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c11_$f.log | head -14
done
echo "== 2. A/B decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 64 > gpurun_out/r2_c11_ab_$name.json 2> gpurun_out/r2_c11_ab_$name.log; echo "$name rc=$? $(cut -c1-330 gpurun_out/r2_c11_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab notail RWKV_B200_NO_LN_TAIL=1
ab noxq RWKV_B200_NO_XQ=1
ab notail_noxq RWKV_B200_NO_LN_TAIL=1 RWKV_B200_NO_XQ=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --out gpurun_out/r2_trace_decode_c11.csv > gpurun_out/r2_trace_decode_c11.log 2>&1; tail -n 44 gpurun_out/r2_trace_decode_c11.log
du -sh gpurun_out
