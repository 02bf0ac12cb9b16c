set -x
O=gpurun_out/r02q
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 120 -x -k conv1x1 > $O/pytest_fused.log 2>&1
tail -25 $O/pytest_fused.log
