set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_gpu.py -x -q 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/gamma_ppf_ulp.json'))
for k,v in d.items(): print(k,v)
PY
timeout 300 python tools/graph_times.py 2>&1 | tail -5
