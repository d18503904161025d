#!/bin/bash
# Round 2, GPU call 18: ln_mix with its mix coefficients loaded before the dependency wait; ncu --set full with source of the prefill
# GEMM (non-split and cluster-split launches) to see where the time after the K loop goes.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
for f in parity; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c18_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c18_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c18_$f.log | head -14
done
echo "== 2. decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 64 > gpurun_out/r2_c18_ab_$name.json 2> gpurun_out/r2_c18_ab_$name.log; echo "$name rc=$? $(cut -c1-330 gpurun_out/r2_c18_ab_$name.json)"; }
ab default RWKV_B200_X=0
echo "== 3. ncu of the prefill GEMM with source"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc" --launch-skip 1 --launch-count 3 -f -o gpurun_out/r2_ncu_gemm_c18 $PY tools/ncu_targets.py > gpurun_out/r2_c18_ncu.log 2>&1; echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
du -sh gpurun_out
