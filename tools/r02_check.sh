#!/bin/bash
# parity tests + memcheck + the headline bench line (pipelined and not).  usage: gpurun -- bash tools/r02_check.sh <tag> [pytest -k filter]
TAG=${1:-chk}; FILTER=${2:-not cfg5}
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q -k "$FILTER" 2>&1 | tail -15) > gpurun_out/${TAG}_tests.log; tail -5 gpurun_out/${TAG}_tests.log
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_frame.py 2>&1 | tail -3)
python bench.py --steps 40 --warmup 5 --no-table --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
OIT_B200_NO_PIPELINE=1 python bench.py --steps 40 --warmup 5 --no-table --no-cpu > gpurun_out/${TAG}_bench_nopipe.json 2>> gpurun_out/${TAG}_bench.err
python - <<EOF
import json
for f in ("${TAG}_bench","${TAG}_bench_nopipe"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, round(d["ms_per_step"],4), d["stages"]["geometry"]["ms"], d["stages"]["clear"]["ms"], d["stages"]["color"]["ms"], d["gpu_launches"]/d["steps"], round(d["e2e"]["ms_per_step"],3), round(d["e2e"]["flattened_mesh"]["ms_per_step"],3))
EOF
tail -3 gpurun_out/${TAG}_bench.err
