set -x
O=gpurun_out/r02b
mkdir -p $O
python -m pytest tests/test_gpu_parity_r2.py tests/test_gpu_gridnet.py tests/test_gpu_corrector.py tests/test_gpu_count_mlp.py -q -m gpu > $O/pytest.log 2>&1
tail -30 $O/pytest.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -3 $O/smoke.log
python bench.py --steps 10 --warmup 3 --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 600 $O/bench_c2.err
python bench.py --config c1 --steps 20 --warmup 3 --profile-out $O/c1_kernels.json > $O/bench_c1.json 2> $O/bench_c1.err; tail -c 600 $O/bench_c1.err
python bench.py --config c4 --steps 5 --warmup 3 --profile-out $O/c4_sweep.json > $O/bench_c4.json 2> $O/bench_c4.err; tail -c 600 $O/bench_c4.err
python bench.py --config c3 --steps 3 --warmup 3 --profile-out $O/c3_kernels.json > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 600 $O/bench_c3.err
python bench.py --config c5 --steps 4 --warmup 3 > $O/bench_c5.json 2> $O/bench_c5.err; tail -c 600 $O/bench_c5.err
python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/bench_nograph.json 2> $O/bench_nograph.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_bf16_kernel|gemm_tn_kernel|conv3x3" -s 1428 -c 357 --csv --log-file $O/dense_traffic.csv python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/ncu1.log 2>&1
python tools/ncu_dense_targets.py > $O/dense_targets_plain.txt 2>&1 && \
ncu --set full --clock-control none -k regex:"gemm_bf16_kernel|gemm_tn_kernel|conv3x3" -c 24 -o /tmp/dense_full python tools/ncu_dense_targets.py > $O/ncu2.log 2>&1
ncu -i /tmp/dense_full.ncu-rep --page raw --csv > $O/dense_full_raw.csv 2>/dev/null
python tools/ncu_targets.py gather > $O/gather_plain.txt 2>&1 && \
ncu --set full --clock-control none -k regex:"patch_gather" -c 4 -o /tmp/gather_full python tools/ncu_targets.py gather > $O/ncu3.log 2>&1
ncu -i /tmp/gather_full.ncu-rep --page raw --csv > $O/gather_full_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la $O
