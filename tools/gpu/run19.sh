set -x
O=gpurun_out/r02s
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_densenet.py -q -m gpu --timeout 200 -x > $O/pytest_f.log 2>&1
tail -3 $O/pytest_f.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out $O/c2_kernels.json > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 300 $O/bench_c2.err
python tools/kernel_table.py $O/c2_kernels.json 20 > $O/kernel_table_c2.txt 2>&1
head -40 $O/kernel_table_c2.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_bf16_kernel|gemm_tn_kernel|conv3x3|gemm_bwd1x1" -s 1196 -c 299 --csv --log-file $O/traffic.csv python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/ncu_traffic.log 2>&1
tail -3 $O/ncu_traffic.log; wc -l $O/traffic.csv
