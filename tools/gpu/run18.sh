set -x
O=gpurun_out/r02r
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > $O/pytest_all.log 2>&1
tail -4 $O/pytest_all.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python tools/kernel_table.py $O/c2_kernels.json 14 > $O/kernel_table_c2.txt 2>&1
head -60 $O/kernel_table_c2.txt
GRIDNEXT_B200_FUSED_BWD1X1=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c2_unfused.json 2> $O/bench_c2_unfused.err
python -c "
import json
for f in ('bench_c2','bench_c2_unfused'):
    d=json.loads(open('$O/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['clocks'])
"
