#!/bin/bash
# per-kernel average durations of a few frames of one workload (ncu launch list, serialised + cold: compare shares)
TAG=${1:-ll}; WL=${2:-headline}
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv python tools/profile_frame.py $WL 6 > /dev/null 2>&1
python - <<EOF
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); ui=hdr.index("Metric Unit")
acc=collections.OrderedDict()
for r in rows[1:]:
    v=float(r[vi].replace(",","")); u=r[ui]
    v = v/1000 if u in ("ns","nsecond") else (v if u in ("us","usecond") else v*1000)
    acc.setdefault(r[ki][:60],[]).append(v)
for k,v in acc.items():
    v=sorted(v); print("%8.1f us median  n=%3d  %s"%(v[len(v)//2],len(v),k))
EOF
