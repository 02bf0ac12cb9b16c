set -x
O=gpurun_out/r02y
mkdir -p $O
timeout 500 python -m pytest tests/test_gpu_gridnet.py tests/test_gpu_count_mlp.py -q -m gpu --timeout 300 -x > $O/pytest_f.log 2>&1
tail -4 $O/pytest_f.log
timeout 400 python bench.py --config c5 --steps 4 --warmup 3 --no-cpu-baseline > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 400 $O/bench_c5.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y/bench_c5.json').read().strip().splitlines()[-1]); print('c5', d['ms_per_step'], d['value'], d['gpu_launches'], d.get('clocks'), d['config'].get('epoch_loss'))
PY
