O=gpurun_out/r02wg
mkdir -p $O
timeout 200 python tools/hexwg_dbg.py > $O/hexwg_dbg.txt 2>&1; cat $O/hexwg_dbg.txt
