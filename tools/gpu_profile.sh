#!/bin/bash
# Profiling half of tools/gpu_check.sh (no tests, no bench line): ncu launch list of the short bench
# command and one `ncu --set full` capture of the sort / post-sort / Gram / transform kernels.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --e2e-steps 0 --cpu-rows 0 > gpurun_out/${TAG}_plain_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --cpu-rows 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list exit $?"
timeout 300 python tools/stage_times.py 2e7 16 2 > gpurun_out/${TAG}_stage_2e7.json 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'partition_pass|post_sort|scatter_rows|col_minmax|sort_hist|gram_small|transform_small' -s 18 -c 18 \
    -o gpurun_out/${TAG}_prof -f python tools/stage_times.py 2e7 16 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv
