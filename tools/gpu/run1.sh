set -x
mkdir -p gpurun_out/r02a
O=gpurun_out/r02a
python bench.py --fused-adam --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_fused.json 2> $O/bench_fused.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/bench_foreach.json 2> $O/bench_foreach.err
python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline > $O/bench_nograph.json 2> $O/bench_nograph.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:gemm_bf16_kernel -s 488 -c 122 --csv --log-file $O/gemm_traffic.csv python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu1.log 2>&1
KB_SPOTS=1248 KB_REPS=1 KB_BLOCKS=13 python tools/kbench.py > $O/kbench_plain.txt 2>&1 && \
KB_SPOTS=1248 KB_REPS=1 KB_BLOCKS=13 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel|gemm_tn_kernel|conv3x3" -c 60 -o $O/dense_full python tools/kbench.py > $O/ncu2.log 2>&1
python tools/ncu_targets.py > $O/targets_plain.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"hexconv|patch_gather|plane|bn_act|masked_ce|bn_stats" -c 80 -o $O/g_full python tools/ncu_targets.py > $O/ncu3.log 2>&1
ls -la $O
