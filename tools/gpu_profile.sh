#!/bin/bash
# Profiling visit (one GPU): ncu launch list of the short bench command and one `ncu --set full` capture of
# every kernel of an Iman-Conover call at the bench size (N=1e8, d=16).  Outputs -> gpurun_out/<tag>_*.
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --e2e-steps 0 --cpu-rows 0 --graph-rows 0 > gpurun_out/${TAG}_plain_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --cpu-rows 0 --graph-rows 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list exit $?"
timeout 300 python tools/stage_times.py 1e8 16 2 > gpurun_out/${TAG}_stage_1e8.json 2>&1 &&
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'pass_tma|post_tma|scatter_rows|col_minmax|sort_hist|gram_small|transform_small' -s 18 -c 14 \
    -o gpurun_out/${TAG}_prof -f python tools/stage_times.py 1e8 16 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
# the report of 1.6e9-key launches is too large to travel: keep its raw page and the digit pass's source page
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:pass_tma \
    --launch-skip 1 --launch-count 1 > gpurun_out/${TAG}_pass_source.csv 2>/dev/null
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:post_tma \
    --launch-skip 0 --launch-count 1 > gpurun_out/${TAG}_post_source.csv 2>/dev/null
ls -la gpurun_out/${TAG}_prof.ncu-rep; rm -f gpurun_out/${TAG}_prof.ncu-rep
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv 33
