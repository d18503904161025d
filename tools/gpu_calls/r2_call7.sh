#!/bin/bash
# Round 2, GPU call 7: diagnose (a) where the prefill GEMM's MMA thread spends its time, (b) the in-kernel LayerNorm of PRO_LN_MIX;
# check the software-pipelined wkv6; ncu of every kernel (sections, all launches) + --set full of the two hot kernels, exported to CSV on the box.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
echo "== 1. parity (wkv6 pipelining)"; timeout 600 $PY -m pytest tests/test_gpu_parity.py tests/test_gpu_batch.py -q -m gpu --timeout 300 --maxfail 12 -rfE > gpurun_out/r2_c7_parity.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/r2_c7_parity.log; grep -E "^(FAILED|ERROR)" gpurun_out/r2_c7_parity.log | head
echo "== 2. prefill trace"
$PY tools/trace_decode.py rwkv6-7b:Q5_1 --prefill 128 --out gpurun_out/r2_trace_prefill_c7.csv > gpurun_out/r2_trace_prefill_c7.log 2>&1; tail -n 42 gpurun_out/r2_trace_prefill_c7.log
echo "== 3. decode: fused LN with marks"
RWKV_B200_FUSE_LN=1 $PY tools/trace_decode.py rwkv6-7b:Q5_1 --out gpurun_out/r2_trace_decode_c7_fuseln.csv > gpurun_out/r2_trace_decode_c7_fuseln.log 2>&1; tail -n 30 gpurun_out/r2_trace_decode_c7_fuseln.log
echo "== 4. ncu"
timeout 420 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --clock-control none -c 140 -f -o /tmp/r2_ncu_sections $PY tools/ncu_targets.py > gpurun_out/r2_c7_ncu_sections.log 2>&1; echo "ncu sections rc=$?"
ncu -i /tmp/r2_ncu_sections.ncu-rep --page raw --csv > gpurun_out/r2_ncu_all_kernels_raw.csv 2>/dev/null; ls -la gpurun_out/r2_ncu_all_kernels_raw.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemv_tma|gemm_tc" -c 12 -f -o /tmp/r2_ncu_hot $PY tools/ncu_targets.py > gpurun_out/r2_c7_ncu_hot.log 2>&1; echo "ncu hot rc=$?"
ncu -i /tmp/r2_ncu_hot.ncu-rep --page raw --csv > gpurun_out/r2_ncu_hot_raw.csv 2>/dev/null; ncu -i /tmp/r2_ncu_hot.ncu-rep --page details > gpurun_out/r2_ncu_hot_details.txt 2>/dev/null; ls -la /tmp/*.ncu-rep gpurun_out/r2_ncu_hot*; 
du -sh gpurun_out
