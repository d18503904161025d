#!/bin/bash
# First GPU call of the next round (run under gpurun from the repo root; ~4 minutes of box time, every step writes to gpurun_out/).
# Validates what round 1 could only write, not run: the experimental bundle RWKV_B200_STAGE_V2=1 (per-block activation staging in
# both decode paths; red.release arrival + L2 prefetches in the persistent kernel) and hunts the persistent-kernel failure at the
# RWKV-5 1.5B shape. Nothing here changes a default.
set -u
mkdir -p gpurun_out
PY=${PY:-python}

# 1. full suite on the defaults (the persistent / experimental checks run last, out of process)
timeout 400 $PY -m pytest tests -q -m gpu --timeout 120 > gpurun_out/r2_suite.log 2>&1; echo "suite rc=$?"

# 2. the experimental staging: per-kernel parity + fixtures must pass unchanged, then what it buys at 7B
RWKV_B200_STAGE_V2=1 timeout 200 $PY -m pytest tests/test_gpu_gemv.py tests/test_gpu_parity.py -q -x -m gpu > gpurun_out/r2_stage_v2_tests.log 2>&1; echo "stage_v2 tests rc=$?"
timeout 120 $PY bench.py --skip-cpu-baseline > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.log; echo "bench rc=$?"
RWKV_B200_STAGE_V2=1 timeout 120 $PY bench.py --skip-cpu-baseline > gpurun_out/r2_bench_stage_v2.json 2> gpurun_out/r2_bench_stage_v2.log; echo "bench v2 rc=$?"
M=/tmp/rwkv_b200_bench/rwkv6-7b-Q5_1-seed1.bin
timeout 60 $PY tools/persistent_probe.py $M --tokens 3 --time 32 --trace > gpurun_out/r2_probe_default.log 2>&1
RWKV_B200_STAGE_V2=1 timeout 60 $PY tools/persistent_probe.py $M --tokens 3 --time 32 --trace > gpurun_out/r2_probe_stage_v2.log 2>&1

# 3. the 1.5B-shape failure under compute-sanitizer (memcheck first; racecheck / synccheck if memcheck is clean)
$PY - <<'PYEOF'
import sys; sys.path.insert(0, "tools")
import synthetic_model as sm
sm.write_direct("/tmp/rwkv5-1b5-Q4_0.bin", "rwkv5-1b5", "Q4_0", seed=3)
PYEOF
timeout 120 $PY tools/persistent_check.py --tokens 4 /tmp/rwkv5-1b5-Q4_0.bin > gpurun_out/r2_1b5_plain.log 2>&1; echo "1b5 persistent only rc=$?"
timeout 120 $PY tools/persistent_check.py --tokens 4 --overlap /tmp/rwkv5-1b5-Q4_0.bin > gpurun_out/r2_1b5_overlap.log 2>&1; echo "1b5 persistent+overlap rc=$?"
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 $PY tools/persistent_check.py --tokens 3 --overlap /tmp/rwkv5-1b5-Q4_0.bin > gpurun_out/r2_1b5_memcheck.log 2>&1; echo "memcheck rc=$?"

# 4. ncu: launch list of one decode step on both paths + one full capture of the persistent kernel
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches.csv $PY bench.py --steps 2 --warmup 1 --skip-cpu-baseline > gpurun_out/r2_ncu_b.log 2>&1
tail -n 5 gpurun_out/r2_suite.log gpurun_out/r2_stage_v2_tests.log gpurun_out/r2_1b5_plain.log gpurun_out/r2_1b5_overlap.log
tail -n 12 gpurun_out/r2_probe_stage_v2.log
