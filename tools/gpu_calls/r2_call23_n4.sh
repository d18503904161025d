#!/bin/bash
# Round 2, 4-GPU call: bench --gpus 4 on the in-library pipeline (peer-memory transport), short.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
timeout 420 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 16 --warmup 3 --prefill-steps 4 > gpurun_out/r2_c23_bench_pp4.json 2> gpurun_out/r2_c23_bench_pp4.log; echo "pp4 rc=$?"; grep -E "pipeline|rank . :|loaded" gpurun_out/r2_c23_bench_pp4.log | tail -8; cut -c1-700 gpurun_out/r2_c23_bench_pp4.json
