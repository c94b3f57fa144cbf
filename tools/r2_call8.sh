set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() { # tag nproc port args...
  tag=$1; np=$2; port=$3; shift 3
  timeout 600 $TR --nproc-per-node $np --master-port $port bench.py --gpus $np "$@" > gpurun_out/r2s_$tag.json 2> gpurun_out/r2s_$tag.err; echo "$tag exit $?"
  python - "$tag" <<'PY'
import json,sys
try:
    d=json.load(open(f'gpurun_out/r2s_{sys.argv[1]}.json'))
    print(sys.argv[1], {k:d[k] for k in ('n_gpus','value','ms_per_step','scaling')}, 'parity', (d.get('parity_check') or {}).get('equal'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
except Exception as e: print('parse fail',e)
PY
}
run weak8 8 29521 --steps 3 --warmup 3 --e2e-steps 2
run strong8 8 29522 --steps 5 --warmup 3 --e2e-steps 0 --scaling strong --parity-rows 0
run weak4 4 29523 --steps 3 --warmup 3 --e2e-steps 2
run strong4 4 29524 --steps 5 --warmup 3 --e2e-steps 0 --scaling strong --parity-rows 0
timeout 300 $TR --nproc-per-node 8 --master-port 29525 tools/host_copy_probe.py 4 > gpurun_out/r2s_probe8.json 2>/dev/null; cat gpurun_out/r2s_probe8.json
timeout 300 $TR --nproc-per-node 8 --master-port 29526 tools/dist_trace.py 1e8 16 > gpurun_out/r2s_trace8.json 2>/dev/null; tail -1 gpurun_out/r2s_trace8.json | cut -c1-400
timeout 300 $TR --nproc-per-node 8 --master-port 29527 tools/dist_trace.py 1.25e7 16 > gpurun_out/r2s_trace8_strong.json 2>/dev/null; tail -1 gpurun_out/r2s_trace8_strong.json | cut -c1-300
