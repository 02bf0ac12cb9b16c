set -x
O=gpurun_out/r02i
mkdir -p $O
timeout 120 python tools/hextc_trace.py > $O/trace_256.txt 2>&1; cat $O/trace_256.txt
B=16 timeout 120 python tools/hextc_trace.py > $O/trace_16.txt 2>&1; cat $O/trace_16.txt
GRIDNEXT_B200_H2_DBG=15 timeout 120 python tools/hextc_trace.py > $O/trace_256_dbg15.txt 2>&1; cat $O/trace_256_dbg15.txt
