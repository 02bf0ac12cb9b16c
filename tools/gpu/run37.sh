O=gpurun_out/r02wg
mkdir -p $O
timeout 300 python bench.py --config c4 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 300 $O/bench_c4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02wg/bench_c4.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value']); print(json.dumps(d['roofline'], indent=0))
PY
timeout 200 python tools/hexwg_dbg.py > $O/hexwg_dbg2.txt 2>&1; grep '"stack": "1"' $O/hexwg_dbg2.txt
