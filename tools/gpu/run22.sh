set -x
O=gpurun_out/r02v
mkdir -p $O
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 900 --csv --log-file $O/launches.csv python bench.py --no-graph --steps 1 --warmup 3 --no-cpu-baseline --no-eager-baseline > $O/ncu_launches.log 2>&1
tail -2 $O/ncu_launches.log | cut -c1-300; wc -l $O/launches.csv
