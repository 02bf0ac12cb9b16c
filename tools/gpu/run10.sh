set -x
O=gpurun_out/r02j
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_corrector.py -q -m gpu --timeout 120 -k "tensor_core" > $O/pytest.log 2>&1
tail -3 $O/pytest.log
timeout 120 python tools/hextc_trace.py > $O/trace_256.txt 2>&1; cat $O/trace_256.txt
timeout 120 python tools/hextc_time.py > $O/hextc_time.txt 2>&1; cat $O/hextc_time.txt
