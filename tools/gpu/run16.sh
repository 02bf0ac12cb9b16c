set -x
O=gpurun_out/r02p
mkdir -p $O
for E in 3 4; do
GN_GEMM_ESTAGES=$E python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-out $O/c2_kernels_e$E.json > $O/bench_c2_e$E.json 2> $O/bench_c2_e$E.err; tail -c 300 $O/bench_c2_e$E.err
python tools/kernel_table.py $O/c2_kernels_e$E.json 40 > $O/kernel_table_e$E.txt 2>&1
done
