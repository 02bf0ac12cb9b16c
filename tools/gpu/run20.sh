set -x
O=gpurun_out/r02t
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_conv.py tests/test_gpu_densenet.py tests/test_gpu_parity_r2.py -q -m gpu --timeout 200 -x > $O/pytest_f.log 2>&1
tail -3 $O/pytest_f.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python tools/kernel_table.py $O/c2_kernels.json 6 > $O/kernel_table_c2.txt 2>&1
head -48 $O/kernel_table_c2.txt
