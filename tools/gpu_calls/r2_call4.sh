#!/bin/bash
# Round 2, GPU call 4: L2 look-ahead micro-benchmark; core tests after the split-K partial-stride fix and with the staged-column
# hand-off (producer kernels emit the GEMV's quantised operand); decode A/B of the hand-off and the per-block staging; prefill A/B.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 0. microbench"; timeout 120 tools/bin/l2_prefetch > gpurun_out/r2_c4_l2_prefetch.txt 2>&1; echo "rc=$?"; cat gpurun_out/r2_c4_l2_prefetch.txt
echo "== 1. smoke"; timeout 300 $PY __graft_entry__.py smoke 2>&1 | tail -n 2
echo "== 2. core tests"
for f in gemv parity pipeline batch overlap; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c4_$f.log 2>&1; echo "$f rc=$?"; tail -n 2 gpurun_out/r2_c4_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c4_$f.log | head -14
done
RWKV_B200_NO_XQ=1 timeout 600 $PY -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c4_parity_noxq.log 2>&1; echo "parity (no hand-off) rc=$?"; tail -n 2 gpurun_out/r2_c4_parity_noxq.log
echo "== 3. A/B decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 48 > gpurun_out/r2_c4_ab_$name.json 2> gpurun_out/r2_c4_ab_$name.log; echo "$name rc=$? $(cut -c1-330 gpurun_out/r2_c4_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab noxq RWKV_B200_NO_XQ=1
ab noxq_v1 RWKV_B200_NO_XQ=1 RWKV_B200_STAGE_V2=0
ab v1 RWKV_B200_STAGE_V2=0
ab fusedecay RWKV_B200_FUSE_DECAY=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --out gpurun_out/r2_trace_decode_c4.csv > gpurun_out/r2_trace_decode_c4.log 2>&1; tail -n 42 gpurun_out/r2_trace_decode_c4.log
echo "== 4. A/B prefill"
pf() { name=$1; shift; env "$@" timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c4_pf_$name.json 2> gpurun_out/r2_c4_pf_$name.log; echo "$name rc=$? $(grep -o 'prefill: [^"]*' gpurun_out/r2_c4_pf_$name.log | tail -1)"; }
pf default RWKV_B200_X=0
pf cs1 RWKV_B200_TC_CLUSTER=1
pf cs2 RWKV_B200_TC_CLUSTER=2
pf nosplit RWKV_B200_TC_SPLITK=1
pf cs1_nosplit RWKV_B200_TC_CLUSTER=1 RWKV_B200_TC_SPLITK=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c4.csv > gpurun_out/r2_trace_prefill_c4.log 2>&1; tail -n 45 gpurun_out/r2_trace_prefill_c4.log
ls gpurun_out | grep c4 | head -60
