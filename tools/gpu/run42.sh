O=gpurun_out/r02fin
mkdir -p $O
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_c2_2gpu.json 2> $O/bench_c2_2gpu.err; tail -c 400 $O/bench_c2_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config c5 --steps 4 --warmup 3 > $O/bench_c5_2gpu.json 2> $O/bench_c5_2gpu.err; tail -c 400 $O/bench_c5_2gpu.err
python - <<'PY'
import json
for f in ('bench_c2_2gpu','bench_c5_2gpu'):
    try:
        d=json.loads(open('gpurun_out/r02fin/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['clocks'])
    except Exception as e: print(f, 'failed', e)
PY
