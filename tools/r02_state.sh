#!/bin/bash
# State of the tree on one B200: all GPU tests, bench lines of every workload, launch list + ncu --set full of the frame kernel.
# usage: gpurun -- bash tools/r02_state.sh <tag>
TAG=${1:-state}
mkdir -p gpurun_out
(time timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/${TAG}_tests.log 2>&1; tail -6 gpurun_out/${TAG}_tests.log
python bench.py --steps 40 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
for k in 1 2 3 4 5; do
  python bench.py --workload config$k --no-table --no-cpu --steps 10 --warmup 3 > gpurun_out/${TAG}_cfg$k.json 2>> gpurun_out/${TAG}_bench.err
done
python - <<EOF2
import json
for f in ["bench"] + ["cfg%d" % k for k in range(1, 6)]:
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % f))
        print(f, round(d["ms_per_step"], 3), round(d["value"] / 1e9, 2), round(d["e2e"]["ms_per_step"], 3), round(d["roofline"]["frac"], 3), d["stages"]["geometry"]["ms"], d["stages"]["color"]["ms"], d.get("parity"))
    except Exception as e:
        print(f, "failed", e)
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print({k: round(v["ms_per_frame"], 3) for k, v in d.get("per_technique_4k_noaa", {}).items()})
EOF2
tail -3 gpurun_out/${TAG}_bench.err
bash tools/launch_list.sh ${TAG} headline
bash tools/r02_prof.sh ${TAG}_prof headline k_raster
