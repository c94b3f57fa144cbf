#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ic_gpu.py tests/test_permcorr_gpu.py -x -q -k "not full_size and not 1e7" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2d_pytest.log
timeout 300 python tools/stage_times.py 1e8 16 3 > gpurun_out/r2d_stage_inter.json 2>&1; echo "interleaved exit $?"
PBL_TICKET_ORDER=column timeout 300 python tools/stage_times.py 1e8 16 2 > gpurun_out/r2d_stage_col.json 2>&1; echo "column exit $?"
grep -h "total_ms\|rank_scores\|rank_gather" gpurun_out/r2d_stage_inter.json gpurun_out/r2d_stage_col.json
timeout 300 python tools/stage_times.py 1e8 16 1 > gpurun_out/r2d_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2d_launches.csv \
    python tools/stage_times.py 1e8 16 1 > gpurun_out/r2d_ncu_launch.log 2>&1
echo "launch list exit $?"
python tools/launch_summary.py gpurun_out/r2d_launches.csv
