#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests.  One tool per invocation (B200_PROFILING.md: running
# several sanitizer tools in one call wedged a GPU on this pool):
#     bash tools/sanitize.sh memcheck|racecheck|synccheck|initcheck [pytest -k expression]
# Default selection: the small-shape cases of every kernel family (hexconv fp32 / tensor-core gen 1 + 2, fused corrector both paths,
# square conv, gather crop + resize, tcgen05 GEMM, 3x3 conv) -- the sanitizer slows kernels 10-50x.
# Writes gpurun_out/sanitizer_<tool>.log; the summary lines are copied into profiles/ by hand after a run.
TOOL=${1:-memcheck}
EXPR=${2:-"shape1 or shape2 or cfg3 or cfg4 or cfg7 or reference_golden or 128-128-64 or 647-7-56 or 333-128-96 or 77-100-504"}
mkdir -p gpurun_out
export GRIDNEXT_B200_SANITIZE=1
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 99 --launch-timeout 120 \
    python -m pytest tests/test_gpu_gemm.py tests/test_gpu_conv.py tests/test_gpu_corrector.py tests/test_gpu_gather.py -q -m gpu -x -k "$EXPR" \
    > gpurun_out/sanitizer_$TOOL.log 2>&1
echo "exit $?" >> gpurun_out/sanitizer_$TOOL.log
grep -E "ERROR SUMMARY|passed|failed|exit " gpurun_out/sanitizer_$TOOL.log | tail -5
