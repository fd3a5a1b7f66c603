#!/bin/bash
TAG=${1:-geo}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_bin_fused|k_sort_scatter" --launch-skip 6 -c 3 -f -o gpurun_out/${TAG} python tools/profile_frame.py headline 4 > gpurun_out/${TAG}.log 2>&1
python tools/ncu_raw_summary.py gpurun_out/${TAG}.ncu-rep > gpurun_out/${TAG}_raw_summary.txt
grep -E "Kernel Name|gpu__time_duration|registers_per_thread|inst_executed.sum |issue_active|warps_active|stalled_(barrier|long|short|wait)" gpurun_out/${TAG}_raw_summary.txt
python tools/ncu_hotlines.py gpurun_out/${TAG}.ncu-rep k_bin_fused 22
python tools/ncu_hotlines.py gpurun_out/${TAG}.ncu-rep k_sort_scatter 14
