O=gpurun_out/r02w
mkdir -p $O
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c2_8gpu.json 2> $O/bench_c2_8gpu.err; echo rc=$?; tail -c 600 $O/bench_c2_8gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02w/bench_c2_8gpu.json').read().strip().splitlines()[-1]); print('c2 x8', d['n_gpus'], d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d.get('clocks'))
except Exception as e: print('ERR', e)
PY
