O=gpurun_out/r02pdl
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu --timeout 300 -x > $O/pytest_all.log 2>&1
tail -3 $O/pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python tools/kernel_table.py $O/c2_kernels.json 4 > $O/kernel_table_c2.txt 2>&1; head -12 $O/kernel_table_c2.txt
GN_NO_PDL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c2_nopdl.json 2> $O/bench_c2_nopdl.err
timeout 200 python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline --no-eager-baseline > $O/bench_c1.json 2> $O/bench_c1.err
GN_NO_PDL=1 timeout 200 python bench.py --config c1 --steps 50 --warmup 5 --no-cpu-baseline --no-eager-baseline > $O/bench_c1_nopdl.json 2> $O/bench_c1_nopdl.err
python - <<'PY'
import json
for f in ('bench_c2','bench_c2_nopdl','bench_c1','bench_c1_nopdl'):
    try:
        d=json.loads(open('gpurun_out/r02pdl/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('last_loss'), (d.get('inference') or {}).get('value'), d['clocks'])
    except Exception as e: print(f, 'failed', e)
PY
