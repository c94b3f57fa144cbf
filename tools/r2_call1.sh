#!/bin/bash
# round 2, first hardware contact of the TMA digit pass + ballot-ranked partition
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_ic_gpu.py -x -q -k "not full_size and not 1e7" > gpurun_out/r2a_pytest_ic.log 2>&1; echo "pytest ic exit $?"
tail -5 gpurun_out/r2a_pytest_ic.log
PBL_PASS_IMPL=classic timeout 300 python tools/stage_times.py 1e8 16 3 > gpurun_out/r2a_stage_classic.json 2>&1; echo "classic exit $?"
timeout 300 python tools/stage_times.py 1e8 16 3 > gpurun_out/r2a_stage_tma.json 2>&1; echo "tma exit $?"
grep -h "total_ms\|rank_scores\|rank_gather" gpurun_out/r2a_stage_classic.json gpurun_out/r2a_stage_tma.json
timeout 600 python bench.py --steps 3 --warmup 3 --e2e-steps 0 --cpu-rows 0 --graph-rows 0 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"
cat gpurun_out/r2a_bench.json
timeout 300 python tools/stage_times.py 1e8 16 1 > gpurun_out/r2a_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2a_launches.csv \
    python tools/stage_times.py 1e8 16 1 > gpurun_out/r2a_ncu_launch.log 2>&1
echo "launch list exit $?"
python tools/launch_summary.py gpurun_out/r2a_launches.csv
