#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "not full_size and not 1e7" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest exit $?"
tail -15 gpurun_out/r2c_pytest.log
timeout 300 python tools/stage_times.py 1e8 16 3 > gpurun_out/r2c_stage_tma.json 2>&1; echo "tma exit $?"
PBL_POST_IMPL=classic timeout 300 python tools/stage_times.py 1e8 16 2 > gpurun_out/r2c_stage_postclassic.json 2>&1; echo "post classic exit $?"
grep -h "total_ms\|rank_scores\|rank_gather" gpurun_out/r2c_stage_tma.json gpurun_out/r2c_stage_postclassic.json
timeout 300 python tools/stage_times.py 1e8 16 1 > gpurun_out/r2c_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2c_launches.csv \
    python tools/stage_times.py 1e8 16 1 > gpurun_out/r2c_ncu_launch.log 2>&1
echo "launch list exit $?"
python tools/launch_summary.py gpurun_out/r2c_launches.csv
