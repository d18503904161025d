#!/bin/bash
# Round 2, GPU call 2: validate + measure what was written while call 1 ran: PRO_LN_MIX / fused decay / L2 look-ahead on the decode
# path, the clustered + split-K tcgen05 GEMM, the in-library pipeline hand-off (one GPU, three stages), v7 / tensor-core batches.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 0. smoke"; timeout 300 $PY __graft_entry__.py smoke 2>&1 | tail -n 2
echo "== 1. per-kernel + parity tests first (fast feedback)"
timeout 900 $PY -m pytest tests/test_gpu_gemv.py tests/test_gpu_parity.py tests/test_gpu_pipeline.py tests/test_gpu_batch.py -q -m gpu --timeout 600 -x > gpurun_out/r2_c2_core.log 2>&1; echo "core rc=$?"; tail -n 15 gpurun_out/r2_c2_core.log
echo "== 2. A/B decode (quick bench: resident decode + timeline)"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 48 > gpurun_out/r2_c2_ab_$name.json 2> gpurun_out/r2_c2_ab_$name.log; echo "$name rc=$? $(cut -c1-260 gpurun_out/r2_c2_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab nofuse_ln RWKV_B200_NO_FUSE_LN=1
ab nofuse_decay RWKV_B200_NO_FUSE_DECAY=1
ab v2 RWKV_B200_STAGE_V2=1
ab xpf RWKV_B200_XPF=1
ab xpf_v2 RWKV_B200_XPF=1 RWKV_B200_STAGE_V2=1
ab pf8 RWKV_B200_L2_PREFETCH=8
ab xpf_pf8_v2 RWKV_B200_XPF=1 RWKV_B200_L2_PREFETCH=8 RWKV_B200_STAGE_V2=1
echo "== 3. A/B prefill"
pf() { name=$1; shift; env "$@" timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c2_pf_$name.json 2> gpurun_out/r2_c2_pf_$name.log; echo "$name rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c2_pf_$name.log | tail -1)"; }
pf default RWKV_B200_X=0
pf cs1 RWKV_B200_TC_CLUSTER=1
pf cs2 RWKV_B200_TC_CLUSTER=2
pf nosplit RWKV_B200_TC_SPLITK=1
pf cs1_nosplit RWKV_B200_TC_CLUSTER=1 RWKV_B200_TC_SPLITK=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c2.csv > gpurun_out/r2_trace_prefill_c2.log 2>&1; tail -n 25 gpurun_out/r2_trace_prefill_c2.log
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --out gpurun_out/r2_trace_decode_c2.csv > gpurun_out/r2_trace_decode_c2.log 2>&1; grep -A12 "critical-path" gpurun_out/r2_trace_decode_c2.log | head -30
echo "== 4. rest of the suite"
timeout 1500 $PY -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_gpu_gemv.py --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_pipeline.py --deselect tests/test_gpu_batch.py -rA > gpurun_out/r2_c2_rest.log 2>&1; echo "rest rc=$?"; tail -n 3 gpurun_out/r2_c2_rest.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c2_rest.log | head -20; grep -E "max\|ours|tensor-core vs" gpurun_out/r2_c2_rest.log | head -20
echo "== 5. batch throughput"
timeout 300 $PY tools/batch_bench.py > gpurun_out/r2_c2_batch_bench.log 2>&1; tail -n 12 gpurun_out/r2_c2_batch_bench.log
echo "== 6. sanitizer on the new kernels"
timeout 900 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 9 $PY -m pytest tests/test_gpu_pipeline.py tests/test_gpu_gemv.py -q -m gpu -x -k "peer_memory or tensor_core or lnmix or matches_oracle" > gpurun_out/r2_c2_memcheck_new.log 2>&1; echo "memcheck rc=$?"; tail -n 4 gpurun_out/r2_c2_memcheck_new.log
timeout 600 compute-sanitizer --tool racecheck --print-limit 30 --error-exitcode 9 $PY -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "long_prompt and 6v0" > gpurun_out/r2_c2_racecheck_tc.log 2>&1; echo "racecheck rc=$?"; tail -n 4 gpurun_out/r2_c2_racecheck_tc.log
ls gpurun_out | grep r2_c2 | head -50
