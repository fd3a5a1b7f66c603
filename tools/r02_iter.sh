#!/bin/bash
# one GPU iteration of round 2: parity tests, sanitizer on small frames, variant timings.  usage: gpurun -- bash tools/r02_iter.sh <tag> [pytest -k filter]
TAG=${1:-iter}; FILTER=${2:-not cfg5}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -k "$FILTER" 2>&1 | tail -15) > gpurun_out/${TAG}_tests.log
tail -4 gpurun_out/${TAG}_tests.log
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_frame.py 2>&1 | tail -8) > gpurun_out/${TAG}_memcheck.log
tail -3 gpurun_out/${TAG}_memcheck.log
timeout 900 python tools/bench_variants.py > gpurun_out/${TAG}_variants.log 2>&1
OIT_B200_LAYERED=1 timeout 300 python tools/bench_variants.py child >> gpurun_out/${TAG}_variants.log 2>&1
cat gpurun_out/${TAG}_variants.log
