O=gpurun_out/r02pf
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_conv.py tests/test_gpu_densenet.py -q -m gpu --timeout 200 -x > $O/pytest_dense.log 2>&1
tail -3 $O/pytest_dense.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python tools/kernel_table.py $O/c2_kernels.json 8 > $O/kernel_table_c2.txt 2>&1; sed -n 1,12p $O/kernel_table_c2.txt; grep -A9 "^gn_conv3x3_bf16\|^gn_conv3x3_wgrad" $O/kernel_table_c2.txt
GN_C3_NO_PREFETCH=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c2_nopf.json 2> $O/bench_c2_nopf.err
python - <<'PY'
import json
for f in ('bench_c2','bench_c2_nopf'):
    try:
        d=json.loads(open('gpurun_out/r02pf/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], (d.get('inference') or {}).get('value'), d['clocks'])
    except Exception as e: print(f, 'failed', e)
PY
