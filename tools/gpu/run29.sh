set -x
O=gpurun_out/r02final
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu --timeout 300 > $O/pytest_all.log 2>&1
tail -4 $O/pytest_all.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -c 300 $O/bench_default.err
timeout 400 python bench.py --impl reference > $O/bench_ref_default.json 2> $O/bench_ref_default.err; tail -c 300 $O/bench_ref_default.err
python - <<'PY'
import json
for f in ('bench_default','bench_ref_default'):
    d=json.loads(open('gpurun_out/r02final/'+f+'.json').read().strip().splitlines()[-1]); print(f, d.get('steps'), d.get('ms_per_step'), d['value'], (d.get('e2e') or {}).get('value'), (d.get('roofline') or {}).get('kernel'), (d.get('roofline') or {}).get('frac'), (d.get('roofline') or {}).get('traffic'), (d.get('cpu_baseline') or {}).get('value'))
PY
