O=gpurun_out/r02xt
mkdir -p $O
for i in 1 2 3; do KB_REPS=15 KB_BLOCKS=1 timeout 120 python tools/kbench.py gemm_xf 2>&1 | grep -E "case|Error" ; done
KB_REPS=9 KB_BLOCKS=234 timeout 120 python tools/kbench.py gemm_xf 2>&1 | grep -E "case|Error"
timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 60 -x 2>&1 | tail -2
