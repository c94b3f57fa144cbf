#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python tools/stage_times.py 2e7 16 2 > gpurun_out/r2e_stage_2e7.json 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pass_tma|post_tma' -s 5 -c 10 \
    -o gpurun_out/r2e_prof -f python tools/stage_times.py 2e7 16 2 > gpurun_out/r2e_ncu_full.log 2>&1
echo "ncu full exit $?"
tail -3 gpurun_out/r2e_ncu_full.log
