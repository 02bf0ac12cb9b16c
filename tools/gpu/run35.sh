O=gpurun_out/r02wg
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_corrector.py -q -m gpu --timeout 120 -x -k "tensor_core" > $O/pytest_tc.log 2>&1
tail -15 $O/pytest_tc.log
timeout 200 python tools/hexwg_time.py > $O/hexwg_time.txt 2>&1; cat $O/hexwg_time.txt
