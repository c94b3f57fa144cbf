#!/bin/bash
# ncu launch list of one stage_times run: bash tools/gpu_launches.sh TAG N K
TAG=$1; N=${2:-1e8}; K=${3:-16}
python tools/stage_times.py $N $K 1 > gpurun_out/${TAG}_stage.json 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python tools/stage_times.py $N $K 1 > gpurun_out/${TAG}_ncu.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv
