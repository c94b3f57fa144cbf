set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ic_gpu.py -x -q -k "not full_size and not 1e7" 2>&1 | tail -3
timeout 600 python tools/stage_times.py 1e6 1024 2 > gpurun_out/r2_config4_stages.json 2>&1; tail -12 gpurun_out/r2_config4_stages.json
timeout 300 python tools/stage_times.py 1e6 1024 1 > /dev/null 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_config4_launches.csv python tools/stage_times.py 1e6 1024 1 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_config4_launches.csv | grep -v "CUDAGen\|init\|scan"
timeout 600 python tools/gamma_truth.py > gpurun_out/r2_gamma_truth.json 2>/dev/null; echo "truth exit $?"
