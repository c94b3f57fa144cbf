set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ic_gpu.py -x -q -k "2_to_the_30 or wide_problem" 2>&1 | tail -3
timeout 600 python tools/stage_times.py 1e6 1024 2 > gpurun_out/r2_config4_stages.json 2>&1; grep -h "gram\|transform\|solve\|total_ms" gpurun_out/r2_config4_stages.json | tail -4
timeout 600 python tools/stage_times.py 2e5 4096 2 2>&1 | grep -h "gram\|transform\|solve\|total_ms" | tail -4
