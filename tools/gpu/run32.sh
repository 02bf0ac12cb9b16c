O=gpurun_out/r02xt
mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_densenet.py -q -m gpu --timeout 100 -x 2>&1 | tail -2
for i in 1 2; do KB_REPS=15 KB_BLOCKS=1234 timeout 200 python tools/kbench.py gemm_xf 2>&1 | grep -E "case|Error" ; done
