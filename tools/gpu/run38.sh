O=gpurun_out/r02wg
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_corrector.py -q -m gpu --timeout 120 -x > $O/pytest_corr.log 2>&1
tail -4 $O/pytest_corr.log
timeout 200 python tools/hextc_time.py > $O/hextc_time.txt 2>&1; grep '"B": 256' $O/hextc_time.txt
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c4b.json 2> $O/bench_c4b.err; tail -c 300 $O/bench_c4b.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02wg/bench_c4b.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value']); print({k:(v['calls'],v['ms']) for k,v in d['roofline']['kernels'].items()})
PY
