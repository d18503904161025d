#!/bin/bash
# Round 2, GPU call 3 (first call of the re-entered session; the outputs of calls 1-2 were lost with the container):
# validate everything written since round 1 and measure it: core tests, decode / prefill A/B, default bench, rest of the suite.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
{ nvidia-smi -L; nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem,memory.total --format=csv; echo "nproc $(nproc)"; cat /sys/fs/cgroup/cpu.max 2>/dev/null; lscpu | egrep "Model name|^CPU\(s\)|Thread|Core|Socket"; free -g | head -2; } > gpurun_out/r2_gpu_box.txt 2>&1
echo "== 0. smoke"; timeout 300 $PY __graft_entry__.py smoke 2>&1 | tail -n 2
echo "== 1. core tests"
for f in gemv parity pipeline batch; do
  timeout 900 $PY -m pytest tests/test_gpu_$f.py -q -m gpu --timeout 300 --maxfail 10 -rfE > gpurun_out/r2_c3_$f.log 2>&1; echo "$f rc=$?"; tail -n 3 gpurun_out/r2_c3_$f.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c3_$f.log | head -12
done
echo "== 2. A/B decode"
ab() { name=$1; shift; env "$@" timeout 200 $PY bench.py --quick --steps 48 > gpurun_out/r2_c3_ab_$name.json 2> gpurun_out/r2_c3_ab_$name.log; echo "$name rc=$? $(cut -c1-300 gpurun_out/r2_c3_ab_$name.json)"; }
ab default RWKV_B200_X=0
ab nofuse_ln RWKV_B200_NO_FUSE_LN=1
ab nofuse_decay RWKV_B200_NO_FUSE_DECAY=1
ab nofuse_both RWKV_B200_NO_FUSE_LN=1 RWKV_B200_NO_FUSE_DECAY=1
ab v2 RWKV_B200_STAGE_V2=1
ab xpf RWKV_B200_XPF=1
ab pf8 RWKV_B200_L2_PREFETCH=8
echo "== 3. A/B prefill"
pf() { name=$1; shift; env "$@" timeout 300 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_c3_pf_$name.json 2> gpurun_out/r2_c3_pf_$name.log; echo "$name rc=$? $(cut -c1-200 gpurun_out/r2_c3_pf_$name.json)"; }
pf default RWKV_B200_X=0
pf cs1 RWKV_B200_TC_CLUSTER=1
pf nosplit RWKV_B200_TC_SPLITK=1
pf cs1_nosplit RWKV_B200_TC_CLUSTER=1 RWKV_B200_TC_SPLITK=1
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c3.csv > gpurun_out/r2_trace_prefill_c3.log 2>&1; tail -n 25 gpurun_out/r2_trace_prefill_c3.log
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --out gpurun_out/r2_trace_decode_c3.csv > gpurun_out/r2_trace_decode_c3.log 2>&1; tail -n 30 gpurun_out/r2_trace_decode_c3.log
echo "== 4. bench"; timeout 600 $PY bench.py > gpurun_out/r2_c3_bench_7b.json 2> gpurun_out/r2_c3_bench_7b.log; echo "bench rc=$?"; tail -n 4 gpurun_out/r2_c3_bench_7b.log; cut -c1-600 gpurun_out/r2_c3_bench_7b.json
echo "== 5. rest of the suite"
timeout 1200 $PY -m pytest tests -q -m gpu --timeout 600 --deselect tests/test_gpu_gemv.py --deselect tests/test_gpu_parity.py --deselect tests/test_gpu_pipeline.py --deselect tests/test_gpu_batch.py -rfE > gpurun_out/r2_c3_rest.log 2>&1; echo "rest rc=$?"; tail -n 3 gpurun_out/r2_c3_rest.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c3_rest.log | head -20
echo "== 6. reference arm"; ( time timeout 300 $PY bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/r2_c3_bench_ref.json 2> gpurun_out/r2_c3_bench_ref.log; echo "ref rc=$?"; tail -n 5 gpurun_out/r2_c3_bench_ref.log; cut -c1-400 gpurun_out/r2_c3_bench_ref.json
ls gpurun_out | head -60
