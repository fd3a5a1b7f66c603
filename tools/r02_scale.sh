#!/bin/bash
# split-frame bench lines on an N-GPU box.  usage: gpurun --gpus 8 -- bash tools/r02_scale.sh <tag> "8 4" [workload] [extra bench args]
TAG=${1:-scale}; NS=${2:-8}; WL=${3:-headline}; EXTRA=${4:-}
mkdir -p gpurun_out
for n in $NS; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 40 --warmup 5 --no-table --no-cpu --workload $WL $EXTRA > gpurun_out/${TAG}_${WL}_n$n.json 2> gpurun_out/${TAG}_${WL}_n$n.err
  python - <<EOF2
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${WL}_n$n.json"))
    print("n=$n", round(d["ms_per_step"], 4), "stages", {k: v["ms"] for k, v in d["stages"].items()}, "e2e", round(d["e2e"]["ms_per_step"], 3), round(d["e2e"]["flattened_mesh"]["ms_per_step"], 3), d["config"]["band_gather_verified"], d.get("exchange_wait"))
except Exception as e:
    print("n=$n failed", e)
EOF2
  tail -2 gpurun_out/${TAG}_${WL}_n$n.err
done
