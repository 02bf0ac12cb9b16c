O=gpurun_out/r02wg
mkdir -p $O
timeout 300 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph > $O/bench_c3_eager.json 2> $O/bench_c3_eager.err; tail -c 300 $O/bench_c3_eager.err
timeout 300 python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_c3_graph.json 2> $O/bench_c3_graph.err; tail -c 300 $O/bench_c3_graph.err
python - <<'PY'
import json
for f in ('bench_c3_eager','bench_c3_graph'):
    try:
        d=json.loads(open('gpurun_out/r02wg/'+f+'.json').read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e'].get('ms_per_step'), d['config'].get('launch'))
    except Exception as e: print(f, 'failed', e)
PY
timeout 120 python tools/ncu_targets.py hexwg > $O/targets_hexwg.txt 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"hexconv_wgrad_tc2|bn_act_bwd" -c 6 -o $O/hexwg_full python tools/ncu_targets.py hexwg > $O/ncu_hexwg.log 2>&1
tail -3 $O/ncu_hexwg.log
ls -la $O
