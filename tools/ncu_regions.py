"""Instruction / stall-sample share of source REGIONS of one kernel in an ncu report:
   python tools/ncu_regions.py rep.ncu-rep k_raster_ll regions.txt
regions.txt: lines `name file first last` (inclusive line ranges); everything else lands in `other:<file>`."""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kern, regf = sys.argv[1], sys.argv[2], sys.argv[3]
regions = []
for ln in open(regf):
    f = ln.split()
    if len(f) == 4 and not ln.startswith("#"):
        regions.append((f[0], f[1], int(f[2]), int(f[3])))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, seen, fn = None, None, None, None
acc = defaultdict(lambda: [0.0, 0.0, 0.0])
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
        seen = seen or fn
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 10 and r[2] == "-" and fn == seen:
        try:
            line, smp, ins, thr = int(r[0]), float(r[hdr["# Samples"]] or 0), float(r[hdr["Instructions Executed"]] or 0), float(r[hdr.get("Thread Instructions Executed", hdr["Instructions Executed"])] or 0)
        except ValueError:
            continue
        name = next((n for n, f, a, b in regions if f == cur and a <= line <= b), f"other:{cur}")
        acc[name][0] += smp
        acc[name][1] += ins
        acc[name][2] += thr
ts, ti, tt = (sum(v[k] for v in acc.values()) or 1 for k in range(3))
print(f"{seen}: samples {ts:.0f}, warp-inst {ti:.0f}, thread-inst {tt:.0f}")
print(f"{'region':<34} {'smp%':>6} {'inst%':>6} {'thr%':>6} {'lanes':>6}")
for n, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:<34} {100*v[0]/ts:6.1f} {100*v[1]/ti:6.1f} {100*v[2]/tt:6.1f} {v[2]/max(v[1],1):6.1f}")
