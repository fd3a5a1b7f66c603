#!/bin/bash
# One-GPU evidence run of a round: bench lines of every workload, the ncu launch list and one --set full capture of the
# frame kernel, all into gpurun_out/ (copy what should be judged into profiles/).  usage: gpurun -- bash tools/final_profile.sh
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r01_bench_final.json 2> gpurun_out/bf.err
for k in 1 2 3 4 5; do
  python bench.py --workload config$k --no-table --no-cpu --steps 10 --warmup 3 > gpurun_out/cfg$k.json 2>> gpurun_out/bf.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-table --no-cpu > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_raster --launch-skip 3 -c 1 -f -o gpurun_out/prof_r01_final2 \
    python tools/profile_frame.py headline 4 > gpurun_out/prof.log 2>&1
tail -2 gpurun_out/prof.log
python - <<'PY'
import json
for f in ["r01_bench_final"] + ["cfg%d" % k for k in range(1, 6)]:
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, round(d["ms_per_step"], 3), round(d["value"] / 1e9, 2), round(d["e2e"]["ms_per_step"], 3), round(d["roofline"]["frac"], 3))
PY
