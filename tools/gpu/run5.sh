set -x
O=gpurun_out/r02e
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_corrector.py tests/test_gpu_parity_r2.py tests/test_gpu_gather.py tests/test_gpu_densenet.py -q -m gpu --timeout 120 > $O/pytest.log 2>&1
tail -8 $O/pytest.log
python tools/kbench_g.py gather > $O/kbench_gather.txt 2>&1; cat $O/kbench_gather.txt
for dbg in 0 1 2 4 8 3 15; do echo "dbg $dbg"; GRIDNEXT_B200_H2_DBG=$dbg timeout 120 python tools/hextc_time.py 2>&1 | grep '"gen": "2"' | grep '"B": 256'; done > $O/hextc_dbg.txt 2>&1
cat $O/hextc_dbg.txt
python bench.py --config c1 --steps 20 --warmup 3 --no-cpu-baseline --profile-out $O/c1_kernels.json > $O/bench_c1.json 2> $O/bench_c1.err; tail -c 300 $O/bench_c1.err
timeout 120 python tools/hextc_time.py > $O/hextc_time.txt 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:"hexconv_tc2_kernel" -s 6 -c 1 -o /tmp/tc2 python tools/hextc_time.py > $O/ncu.log 2>&1
ncu -i /tmp/tc2.ncu-rep --page raw --csv > $O/tc2_raw.csv 2>/dev/null
ncu -i /tmp/tc2.ncu-rep --page source --csv > $O/tc2_source.csv 2>/dev/null
ls -la /tmp/tc2.ncu-rep
python tools/ncu_targets.py gather > $O/targets_plain.txt 2>&1 && \
ncu --set full --clock-control none -k regex:"patch_gather" -c 2 -o /tmp/gat python tools/ncu_targets.py gather > $O/ncu2.log 2>&1
ncu -i /tmp/gat.ncu-rep --page raw --csv > $O/gather_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la $O
