#!/bin/bash
# Round 2, 2-GPU call: the in-library layer pipeline on real peer memory: pipeline tests + bench --gpus 2 (with its 1-GPU logits check).
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
nvidia-smi -L; nvidia-smi topo -m | head -8
echo "== 1. pipeline tests (2 GPUs visible)"; timeout 600 $PY -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -rfEs > gpurun_out/r2_c8_pipeline_n2.log 2>&1; echo "rc=$?"; tail -n 4 gpurun_out/r2_c8_pipeline_n2.log
echo "== 2. bench --gpus 2 (Q5_1)"
timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/r2_c8_bench_pp2.json 2> gpurun_out/r2_c8_bench_pp2.log; echo "pp2 rc=$?"; tail -n 6 gpurun_out/r2_c8_bench_pp2.log; cut -c1-900 gpurun_out/r2_c8_bench_pp2.json
echo "== 3. bench --gpus 2 (config 4: Q8_0)"
timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 4 --workload rwkv6-7b:Q8_0 > gpurun_out/r2_c8_bench_pp2_q8.json 2> gpurun_out/r2_c8_bench_pp2_q8.log; echo "pp2 q8 rc=$?"; tail -n 3 gpurun_out/r2_c8_bench_pp2_q8.log; cut -c1-500 gpurun_out/r2_c8_bench_pp2_q8.json
echo "== 4. in-process pipeline behind rwkv.h (RWKV_B200_PIPELINE_DEVICES=0,1)"
RWKV_B200_PIPELINE_DEVICES=0,1 timeout 600 $PY -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "fixture_logits or chunked_equals or long_prompt" --maxfail 5 -rfE > gpurun_out/r2_c8_parity_pipe2.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2_c8_parity_pipe2.log
