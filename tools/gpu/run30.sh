O=gpurun_out/r02xt
mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_gemm.py -q -m gpu --timeout 60 -x -k "tensor_memory or operand_transform" > $O/pytest_xt.log 2>&1
tail -5 $O/pytest_xt.log
KB_BLOCKS=12 timeout 120 python tools/kbench.py gemm_xf > $O/kb_xt.txt 2>&1; cat $O/kb_xt.txt
GN_GEMM_NO_XT=1 KB_BLOCKS=12 timeout 120 python tools/kbench.py gemm_xf > $O/kb_noxt.txt 2>&1; cat $O/kb_noxt.txt
