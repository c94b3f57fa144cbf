#!/bin/bash
# One GPU-box visit: parity tests, the default bench, an ncu launch list of the bench command and
# one `ncu --set full` capture of the sort digit pass + post-sort kernels.  Outputs -> gpurun_out/.
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/${TAG}_pytest.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --steps 1 --warmup 3 --e2e-steps 0 --cpu-rows 0 > gpurun_out/${TAG}_plain_short.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 0 --cpu-rows 0 > gpurun_out/${TAG}_ncu_launch.log 2>&1
echo "launch list exit $?"
timeout 300 python tools/stage_times.py 2e7 16 2 > gpurun_out/${TAG}_stage_2e7.json 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pass_tma|post_tma|scatter_rows|col_minmax|sort_hist|gram_small|transform_small' -s 18 -c 18 \
    -o gpurun_out/${TAG}_prof -f python tools/stage_times.py 2e7 16 2 > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i gpurun_out/${TAG}_prof.ncu-rep --page raw --csv > gpurun_out/${TAG}_prof_raw.csv 2>/dev/null
tail -3 gpurun_out/${TAG}_pytest.log; cat gpurun_out/${TAG}_bench.json
