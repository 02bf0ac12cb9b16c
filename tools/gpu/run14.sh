set -x
O=gpurun_out/r02n
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x > $O/pytest_all.log 2>&1
tail -6 $O/pytest_all.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
python bench.py --steps 10 --warmup 3 --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python tools/kernel_table.py $O/c2_kernels.json > $O/kernel_table_c2.txt 2>&1
head -30 $O/kernel_table_c2.txt
ls -la $O
