set -u
mkdir -p gpurun_out
O=gpurun_out/r2w
timeout 900 python -m pytest tests/test_ic_gpu.py -x -q -k "one_tile_per_block or random_problems or golden or c_order" 2>&1 | tail -2
st() { # tag post n k cb
  PBL_POST_IMPL=$2 timeout 300 python tools/stage_times.py $3 $4 2 $5 > ${O}_stage_$1.json 2>&1
  python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r2w_stage_{sys.argv[1]}.json'))
    r=d['reps'][-1]; print(sys.argv[1], {k:round(v,2) for k,v in r.items() if k in('rank_scores','rank_gather','total_ms')})
except Exception as e: print(sys.argv[1],'parse fail',e)
PY
}
st n2e8_tma tma 2e8 2 1
st n2e8_classic classic 2e8 2 1
st n4e8_tma tma 4e8 2 1
st n4e8_classic classic 4e8 2 1
st n8e8_auto "" 8e8 2 1
timeout 300 python tools/graph_times.py 2e7 > ${O}_graph_2e7.json 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:graph_eval -c 4 -o ${O}_graph_prof -f python tools/graph_times.py 2e7 > ${O}_graph_ncu.log 2>&1
echo "ncu exit $?"
ncu -i ${O}_graph_prof.ncu-rep --page raw --csv > ${O}_graph_raw.csv 2>/dev/null
ncu -i ${O}_graph_prof.ncu-rep --page source --csv > ${O}_graph_source.csv 2>/dev/null
rm -f ${O}_graph_prof.ncu-rep
ls -la gpurun_out | grep r2w | head
