O=gpurun_out/r02wg
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_corrector.py -q -m gpu --timeout 120 -x -k "tensor_core" > $O/pytest_tc.log 2>&1
tail -3 $O/pytest_tc.log
timeout 200 python tools/hextc_time.py > $O/hextc_time.txt 2>&1; grep '"B": 256' $O/hextc_time.txt
