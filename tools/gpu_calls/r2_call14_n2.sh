#!/bin/bash
# Round 2, 2-GPU call: the in-library layer pipeline on real peer memory: pipeline tests + bench --gpus 2 (with its 1-GPU logits check),
# the in-process pipeline behind rwkv.h.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
nvidia-smi -L; nvidia-smi topo -m | head -6
echo "== 1. pipeline tests (2 GPUs visible)"; timeout 600 $PY -m pytest tests/test_gpu_pipeline.py -q -m gpu --timeout 300 -rfEs > gpurun_out/r2_c14_pipeline_n2.log 2>&1; echo "rc=$?"; tail -n 6 gpurun_out/r2_c14_pipeline_n2.log
echo "== 2. bench --gpus 2 (Q5_1)"
timeout 900 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 4 > gpurun_out/r2_c14_bench_pp2.json 2> gpurun_out/r2_c14_bench_pp2.log; echo "pp2 rc=$?"; tail -n 8 gpurun_out/r2_c14_bench_pp2.log; cut -c1-1200 gpurun_out/r2_c14_bench_pp2.json
echo "== 3. reference arm under torchrun"
timeout 600 $PY -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 4 --warmup 1 > gpurun_out/r2_c14_ref_pp2.json 2> gpurun_out/r2_c14_ref_pp2.log; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2_c14_ref_pp2.json
echo "== 4. in-process pipeline behind rwkv.h (RWKV_B200_PIPELINE_DEVICES=0,1)"
RWKV_B200_PIPELINE_DEVICES=0,1 timeout 600 $PY -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "fixture_logits or chunked_equals or long_prompt" --maxfail 5 -rfE > gpurun_out/r2_c14_parity_pipe2.log 2>&1; echo "rc=$?"; tail -n 3 gpurun_out/r2_c14_parity_pipe2.log
du -sh gpurun_out
