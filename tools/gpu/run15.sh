set -x
O=gpurun_out/r02o
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_gather.py -q -m gpu --timeout 300 -x > $O/pytest_gather.log 2>&1
tail -3 $O/pytest_gather.log
python tools/kbench_g.py gather > $O/kbench_gather.txt 2>&1
cat $O/kbench_gather.txt
