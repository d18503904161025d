#!/bin/bash
# Round 2, GPU call 22: graded layer groups for the overlapped host-state copies (e2e), A/B against the even split.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. tests"
timeout 600 $PY -m pytest tests/test_gpu_overlap.py -q -m gpu --timeout 300 -rfE > gpurun_out/r2_c22_overlap.log 2>&1; echo "overlap rc=$?"; tail -n 2 gpurun_out/r2_c22_overlap.log
timeout 900 $PY -m pytest tests/test_gpu_full_shape.py -q -m gpu --timeout 600 -k "decode_matches or invariants" -rfE > gpurun_out/r2_c22_full_shape.log 2>&1; echo "full shape rc=$?"; tail -n 3 gpurun_out/r2_c22_full_shape.log
echo "== 2. bench e2e: graded vs even groups"
timeout 400 $PY bench.py --steps 64 --skip-cpu-baseline > gpurun_out/r2_c22_bench_graded.json 2> gpurun_out/r2_c22_bench_graded.log; echo "rc=$?"; grep -E "decode|e2e" gpurun_out/r2_c22_bench_graded.log | tail -4
RWKV_B200_EVEN_SEGMENTS=1 timeout 400 $PY bench.py --steps 64 --skip-cpu-baseline > gpurun_out/r2_c22_bench_even.json 2> gpurun_out/r2_c22_bench_even.log; echo "rc=$?"; grep -E "decode|e2e" gpurun_out/r2_c22_bench_even.log | tail -4
du -sh gpurun_out
