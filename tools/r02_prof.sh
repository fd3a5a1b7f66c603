#!/bin/bash
# ncu --set full capture of the frame kernel of one workload + digests.  usage: gpurun -- bash tools/r02_prof.sh <tag> [workload] [kernel regex]
TAG=${1:-prof}; WL=${2:-headline}; KRE=${3:-k_raster}
mkdir -p gpurun_out
ncu --set full --metrics lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_op_atom.sum.per_second,l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum,l1tex__t_set_accesses_pipe_lsu_mem_global_op_red.sum,smsp__inst_executed_op_global_atom.sum --clock-control none --import-source on -k regex:$KRE --launch-skip 3 -c 1 -f -o gpurun_out/${TAG} \
    python tools/profile_frame.py $WL 4 > gpurun_out/${TAG}.log 2>&1
tail -2 gpurun_out/${TAG}.log
python tools/ncu_raw_summary.py gpurun_out/${TAG}.ncu-rep > gpurun_out/${TAG}_raw_summary.txt
python tools/ncu_hotlines.py gpurun_out/${TAG}.ncu-rep $KRE 60 > gpurun_out/${TAG}_hotlines.txt 2>&1
python tools/make_traffic.py gpurun_out/${TAG}.ncu-rep $WL gpurun_out/${TAG}_traffic.json
head -40 gpurun_out/${TAG}_raw_summary.txt
