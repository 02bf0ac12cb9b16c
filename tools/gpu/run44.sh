O=gpurun_out/r02fin
mkdir -p $O
timeout 170 python -m pytest tests -q -m gpu --timeout 120 -x > $O/pytest_all.log 2>&1
tail -3 $O/pytest_all.log
timeout 40 python __graft_entry__.py smoke > $O/smoke.log 2>&1; tail -2 $O/smoke.log
