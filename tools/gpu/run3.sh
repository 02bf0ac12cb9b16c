set -x
O=gpurun_out/r02c
mkdir -p $O
python -m pytest tests/test_gpu_corrector.py tests/test_gpu_parity_r2.py tests/test_gpu_gather.py tests/test_gpu_gridnet.py -q -m gpu > $O/pytest.log 2>&1
tail -25 $O/pytest.log
python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -3 $O/smoke.log
python bench.py --config c1 --steps 20 --warmup 3 --profile-out $O/c1_kernels.json > $O/bench_c1.json 2> $O/bench_c1.err; tail -c 400 $O/bench_c1.err
python bench.py --config c3 --steps 3 --warmup 3 --profile-out $O/c3_kernels.json > $O/bench_c3.json 2> $O/bench_c3.err; tail -c 600 $O/bench_c3.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 400 $O/bench_c2.err
python tools/kbench_g.py corrector > $O/kbench_corrector.txt 2>&1
du -sh gpurun_out; ls -la $O
