#!/bin/bash
# Round 2, last GPU call: the whole -m gpu suite and the default bench line on the final commit.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
timeout 400 $PY -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/r2_last_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/r2_last_gpu_tests.log
timeout 200 $PY bench.py --cpu-budget-s 8 > gpurun_out/r2_last_bench_7b.json 2> gpurun_out/r2_last_bench_7b.log; echo "bench rc=$?"; grep -E "decode|prefill" gpurun_out/r2_last_bench_7b.log | tail -5; cut -c1-200 gpurun_out/r2_last_bench_7b.json
