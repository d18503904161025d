#!/bin/bash
# Round 2, final evidence call: the whole GPU suite as the driver runs it, smoke(), the bench lines (both arms, every BASELINE config),
# the ncu launch list of the bench command, ncu of every kernel + --set full of the hot ones, DRAM traffic of the dominant launch.
set -u
mkdir -p gpurun_out
PY=${PY:-python}
export RWKV_B200_BENCH_DIR=/tmp/rwkv_b200_bench
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/r2_final_gpu_box.txt; nproc >> gpurun_out/r2_final_gpu_box.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_final_gpu_box.txt 2>/dev/null; lscpu | head -20 >> gpurun_out/r2_final_gpu_box.txt
echo "== 1. pytest -m gpu (whole suite)"
timeout 1500 $PY -m pytest tests -q -m gpu --timeout 600 -rfEs > gpurun_out/r2_final_gpu_tests.log 2>&1; echo "rc=$?"; tail -n 4 gpurun_out/r2_final_gpu_tests.log
echo "== 2. smoke"
timeout 300 $PY -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1; echo "rc=$?"; tail -n 2 gpurun_out/r2_final_smoke.log
echo "== 3. bench: ours + reference arm"
timeout 600 $PY bench.py --cpu-budget-s 12 > gpurun_out/r2_final_bench_7b.json 2> gpurun_out/r2_final_bench_7b.log; echo "ours rc=$?"; grep -E "decode|prefill|cpu" gpurun_out/r2_final_bench_7b.log | tail -8
timeout 600 $PY bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r2_final_bench_7b_reference.json 2> gpurun_out/r2_final_bench_7b_reference.log; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2_final_bench_7b_reference.json
timeout 400 $PY bench.py --mode prefill --steps 12 --skip-cpu-baseline > gpurun_out/r2_final_bench_7b_prefill.json 2> gpurun_out/r2_final_bench_7b_prefill.log; echo "prefill rc=$?"
echo "== 4. other BASELINE configs"
timeout 300 $PY bench.py --workload rwkv4-169m:Q5_1 --steps 256 --cpu-budget-s 6 > gpurun_out/r2_final_bench_169m.json 2> gpurun_out/r2_final_bench_169m.log; echo "169m rc=$? $(cut -c1-120 gpurun_out/r2_final_bench_169m.json)"
timeout 400 $PY bench.py --workload rwkv5-1b5:Q4_0 --mode prefill --steps 16 --cpu-budget-s 10 > gpurun_out/r2_final_bench_1b5_prefill.json 2> gpurun_out/r2_final_bench_1b5_prefill.log; echo "1b5 rc=$?"
timeout 400 $PY bench.py --workload rwkv7-2b9:FP16 --cpu-budget-s 10 > gpurun_out/r2_final_bench_2b9.json 2> gpurun_out/r2_final_bench_2b9.log; echo "2b9 rc=$?"
echo "== 5. ncu"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_final_ncu_launches_bench.csv $PY bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r2_final_ncu_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 420 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --section WarpStateStats --section SchedulerStats --clock-control none -c 160 -f -o /tmp/r2_final_sections $PY tools/ncu_targets.py > gpurun_out/r2_final_ncu_sections.log 2>&1; echo "ncu sections rc=$?"
ncu -i /tmp/r2_final_sections.ncu-rep --page raw --csv > gpurun_out/r2_final_ncu_all_kernels_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemv_tma|gemm_tc" -c 14 -f -o /tmp/r2_final_hot $PY tools/ncu_targets.py > gpurun_out/r2_final_ncu_hot.log 2>&1; echo "ncu hot rc=$?"
ncu -i /tmp/r2_final_hot.ncu-rep --page raw --csv > gpurun_out/r2_final_ncu_hot_raw.csv 2>/dev/null; ncu -i /tmp/r2_final_hot.ncu-rep --page details > gpurun_out/r2_final_ncu_hot_details.txt 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"v6_lerp|wkv6|wkv7|wkv4|convert_f16|ln_mix|embed_ln0|sample_kernel" -c 16 -f -o /tmp/r2_final_glue $PY tools/ncu_targets.py > gpurun_out/r2_final_ncu_glue.log 2>&1; echo "ncu glue rc=$?"
ncu -i /tmp/r2_final_glue.ncu-rep --page raw --csv > gpurun_out/r2_final_ncu_glue_raw.csv 2>/dev/null; ncu -i /tmp/r2_final_glue.ncu-rep --page details > gpurun_out/r2_final_ncu_glue_details.txt 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:"gemv_tma" --launch-skip 3 --launch-count 5 -f -o gpurun_out/r2_final_ncu_gemv_7b $PY bench.py --quick --steps 1 > gpurun_out/r2_final_ncu_gemv_7b.log 2>&1; echo "ncu gemv 7b rc=$?"
ls -la gpurun_out/*.ncu-rep /tmp/*.ncu-rep 2>/dev/null
du -sh gpurun_out
