set -x
O=gpurun_out/r02g
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_corrector.py tests/test_gpu_gather.py -q -m gpu --timeout 120 -k "tensor_core or gather" > $O/pytest.log 2>&1
tail -4 $O/pytest.log
python tools/kbench_g.py gather > $O/kbench_gather.txt 2>&1; cat $O/kbench_gather.txt
for dbg in 0 15; do echo "dbg $dbg"; GRIDNEXT_B200_H2_DBG=$dbg timeout 120 python tools/hextc_time.py 2>&1 | grep '"gen": "2"'; done > $O/hextc_dbg.txt 2>&1
grep -v "^+" $O/hextc_dbg.txt
bash tools/sanitize.sh memcheck
cp gpurun_out/sanitizer_memcheck.log $O/
