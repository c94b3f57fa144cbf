set -u
mkdir -p gpurun_out
O=gpurun_out/r2u
PBL_LOOKBACK=narrow timeout 600 python -m pytest tests/test_ic_gpu.py -x -q -k "random_problems or heavy_ties or window or chunk_hook or c_order or golden or dense" 2>&1 | tail -2
timeout 900 python -m pytest tests/test_ic_gpu.py -x -q 2>&1 | tail -2
st() { # tag lb n k cb
  PBL_LOOKBACK=$2 timeout 300 python tools/stage_times.py $3 $4 2 $5 > ${O}_stage_$1.json 2>&1
  python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r2u_stage_{sys.argv[1]}.json'))
    r=d['reps'][-1]; print(sys.argv[1], {k:round(v,2) for k,v in r.items() if k in('rank_scores','rank_gather','total_ms')})
except Exception as e: print(sys.argv[1],'parse fail',e)
PY
}
st base16 auto 1e8 16 0
st base16_narrow narrow 1e8 16 0
st k2_wide wide 1e8 2 1
st k2_narrow narrow 1e8 2 1
st n8e8_wide wide 8e8 2 1
st n8e8_narrow narrow 8e8 2 1
for lb in wide narrow; do PBL_LIB=$PWD/probabilit_b200/libpbl_stats.so PBL_LOOKBACK=$lb timeout 300 python tools/tile_stats.py 1e8 2 1 2>&1 | tail -1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches_8e8.csv python tools/stage_times.py 8e8 2 1 1 > ${O}_ncu.log 2>&1; echo "ncu exit $?"
python tools/launch_summary.py ${O}_launches_8e8.csv 2>&1 | tail -16
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches_1e8.csv python tools/stage_times.py 1e8 16 1 0 > ${O}_ncu2.log 2>&1; echo "ncu exit $?"
python tools/launch_summary.py ${O}_launches_1e8.csv 2>&1 | tail -16
