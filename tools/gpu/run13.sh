set -x
O=gpurun_out/r02m
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --timeout 300 > $O/pytest_all.log 2>&1
tail -6 $O/pytest_all.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py --steps 10 --warmup 3 --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python bench.py --config c1 --steps 20 --warmup 3 --profile-out $O/c1_kernels.json > $O/bench_c1.json 2> $O/bench_c1.err; tail -c 300 $O/bench_c1.err
python bench.py --config c4 --steps 5 --warmup 3 --profile-out $O/c4_sweep.json > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 300 $O/bench_c4.err
python bench.py --config c3 --steps 3 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 300 $O/bench_c3.err
python bench.py --config c5 --steps 4 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 300 $O/bench_c5.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
python tools/kbench_g.py gather corrector ce mlp > $O/kbench_g.txt 2>&1
ls -la $O
