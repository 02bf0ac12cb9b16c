set -x
O=gpurun_out/r02d
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_corrector.py tests/test_gpu_parity_r2.py tests/test_gpu_gather.py -q -m gpu -k "not tensor_core" --timeout 120 > $O/pytest.log 2>&1
tail -12 $O/pytest.log
python bench.py --config c1 --steps 20 --warmup 3 --no-cpu-baseline --profile-out $O/c1_kernels.json > $O/bench_c1.json 2> $O/bench_c1.err; tail -c 400 $O/bench_c1.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err; tail -c 400 $O/bench_c2.err
python tools/kbench_g.py gather corrector > $O/kbench_g.txt 2>&1
python tools/ncu_targets.py gather corrector > $O/targets_plain.txt 2>&1 && \
ncu --set full --clock-control none -k regex:"patch_gather|corrector_fused" -c 12 -o /tmp/g_full python tools/ncu_targets.py gather corrector > $O/ncu.log 2>&1
ncu -i /tmp/g_full.ncu-rep --page raw --csv > $O/g_full_raw.csv 2>/dev/null
# new tensor-core hexconv (second generation): its own process, bounded
timeout 300 python -m pytest tests/test_gpu_corrector.py -q -m gpu -k "tensor_core" --timeout 60 > $O/pytest_tc.log 2>&1
tail -15 $O/pytest_tc.log
GRIDNEXT_B200_G_FUSED=0 timeout 120 python tools/hextc_time.py > $O/hextc_time.txt 2>&1; tail -5 $O/hextc_time.txt
du -sh gpurun_out; ls -la $O
