set -x
O=gpurun_out/r02u
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu --timeout 300 > $O/pytest_all.log 2>&1
tail -4 $O/pytest_all.log
timeout 200 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python tools/kernel_table.py $O/c2_kernels.json 30 > $O/kernel_table_c2.txt 2>&1
timeout 300 python bench.py --config c1 --steps 20 --warmup 3 --profile-out $O/c1_kernels.json > $O/bench_c1.json 2> $O/bench_c1.err; tail -c 300 $O/bench_c1.err
timeout 400 python bench.py --config c4 --steps 5 --warmup 3 --profile-out $O/c4_sweep.json > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 300 $O/bench_c4.err
timeout 400 python bench.py --config c3 --steps 3 --warmup 3 > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 300 $O/bench_c3.err
timeout 400 python bench.py --config c5 --steps 4 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 300 $O/bench_c5.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 420 --csv --log-file $O/launches.csv python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none -k regex:"gemm_bf16_kernel|gemm_tn_kernel|conv3x3|gemm_bwd1x1" -c 28 -o /tmp/dense_full python tools/ncu_dense_targets.py > $O/ncu2.log 2>&1
ncu -i /tmp/dense_full.ncu-rep --page raw --csv > $O/dense_full_raw.csv 2>/dev/null
timeout 300 python tools/kbench_g.py gather corrector ce mlp > $O/kbench_g.txt 2>&1
ls -la $O
