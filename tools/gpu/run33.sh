O=gpurun_out/r02xt
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu --timeout 300 > $O/pytest_all.log 2>&1
tail -3 $O/pytest_all.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python tools/kernel_table.py $O/c2_kernels.json 4 > $O/kernel_table_c2.txt 2>&1; head -12 $O/kernel_table_c2.txt
GN_GEMM_NO_XT=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c2_noxt.json 2> $O/bench_c2_noxt.err
python - <<'PY'
import json
for f in ('bench_c2','bench_c2_noxt'):
    d=json.loads(open('gpurun_out/r02xt/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['inference']['value'], d['clocks'])
PY
