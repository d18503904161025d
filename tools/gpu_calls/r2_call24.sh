#!/bin/bash
# Round 2, GPU call 24: ln_mix with all (up to 6) coefficient vectors requested before the dependency wait for v4 / v5.2 / v7 blocks.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
timeout 600 $PY -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c24_parity.log 2>&1; echo "parity+batch rc=$?"; tail -n 2 gpurun_out/r2_c24_parity.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c24_parity.log | head
timeout 300 $PY bench.py --workload rwkv7-2b9:FP16 --quick --steps 64 > gpurun_out/r2_c24_ab_2b9.json 2> gpurun_out/r2_c24_ab_2b9.log; echo "2b9 rc=$? $(cut -c1-300 gpurun_out/r2_c24_ab_2b9.json)"
timeout 200 $PY bench.py --quick --steps 64 > gpurun_out/r2_c24_ab_7b.json 2> gpurun_out/r2_c24_ab_7b.log; echo "7b rc=$? $(cut -c1-300 gpurun_out/r2_c24_ab_7b.json)"
