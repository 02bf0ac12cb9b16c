O=gpurun_out/r02k
mkdir -p $O
timeout 120 python tools/hextc_trace.py > $O/trace_256.txt 2>&1; cat $O/trace_256.txt
GRIDNEXT_B200_H2_DBG=1 timeout 120 python tools/hextc_trace.py > $O/trace_256_nostore.txt 2>&1; cat $O/trace_256_nostore.txt
