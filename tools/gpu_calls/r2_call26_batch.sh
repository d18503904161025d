#!/bin/bash
# Round 2: aggregate throughput of batch contexts at the 7B shape on the final kernels (tools/batch_bench.py).
mkdir -p gpurun_out
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
timeout 150 python tools/batch_bench.py > gpurun_out/r2_c26_batch_bench.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2_c26_batch_bench.txt | tail -9
