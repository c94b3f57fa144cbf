set -u
mkdir -p gpurun_out
O=gpurun_out/r2v
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_ic_gpu.py -x -q -k "not full_size and not 1e7 and not 2_to_the_30" 2>&1 | tail -3
for r in 0 2 4; do PBL_GRAPH_ROWS=$r timeout 300 python tools/graph_times.py 1e8 > ${O}_graph_r$r.json 2>&1; python - $r <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r2v_graph_r{sys.argv[1]}.json'))
    print('rows',sys.argv[1],{k:(round(min(v.get('wall_s_incl_d2h_of_sink',v.get('wall_s_kernel_only')))*1e3,1)) for k,v in d.items() if isinstance(v,dict)})
except Exception as e: print('parse fail',e); print(open(f'gpurun_out/r2v_graph_r{sys.argv[1]}.json').read()[-600:])
PY
done
st() { # tag pass post n k cb
  PBL_PASS_IMPL=$2 PBL_POST_IMPL=$3 timeout 300 python tools/stage_times.py $4 $5 2 $6 > ${O}_stage_$1.json 2>&1
  python - "$1" <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r2v_stage_{sys.argv[1]}.json'))
    r=d['reps'][-1]; print(sys.argv[1], {k:round(v,2) for k,v in r.items() if k in('rank_scores','rank_gather','total_ms')})
except Exception as e: print(sys.argv[1],'parse fail',e)
PY
}
st base16 tma tma 1e8 16 0
st k2 tma tma 1e8 2 1
st k2_cpass classic tma 1e8 2 1
st k2_cboth classic classic 1e8 2 1
st n8e8 tma tma 8e8 2 1
st n8e8_cpass classic tma 8e8 2 1
st n8e8_cboth classic classic 8e8 2 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches_8e8.csv python tools/stage_times.py 8e8 2 1 1 > ${O}_ncu.log 2>&1; echo "ncu exit $?"
python tools/launch_summary.py ${O}_launches_8e8.csv 2>&1 | grep "pass_tma\|post_tma\|scatter"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_launches_1e8.csv python tools/stage_times.py 1e8 16 1 0 > ${O}_ncu2.log 2>&1; echo "ncu exit $?"
python tools/launch_summary.py ${O}_launches_1e8.csv 2>&1 | grep "pass_tma\|post_tma\|scatter"
