O=gpurun_out/r02fin
mkdir -p $O
timeout 110 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c2_2gpu.json 2> $O/bench_c2_2gpu.err; echo rc=$?; tail -c 300 $O/bench_c2_2gpu.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02fin/bench_c2_2gpu.json').read().strip().splitlines()[-1]); print('c2 x2', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
except Exception as e: print('failed', e)
PY
