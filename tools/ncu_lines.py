"""Per-line instruction / sample shares of one source file of a kernel, in line order (developer tool):
python tools/ncu_lines.py rep.ncu-rep k_raster oit_raster.cu [min_inst_pct]"""
import csv
import subprocess
import sys

rep, kern, want = sys.argv[1], sys.argv[2], sys.argv[3]
thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.15
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, data, first, fn = None, None, [], None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        first = first or r[1]
        fn = r[1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 10 and r[2] == "-" and fn == first:
        try:
            data.append((cur, int(r[0]), r[1].strip()[:110], float(r[hdr["# Samples"]] or 0), float(r[hdr["Instructions Executed"]] or 0),
                         float(r[hdr["Thread Instructions Executed"]] or 0) if "Thread Instructions Executed" in hdr else 0.0))
        except ValueError:
            pass
ts, ti = sum(d[3] for d in data) or 1, sum(d[4] for d in data) or 1
files = {}
for d in data:
    f = files.setdefault(d[0], [0.0, 0.0])
    f[0] += d[3]
    f[1] += d[4]
for f, (s, i) in sorted(files.items(), key=lambda kv: -kv[1][1]):
    print(f"{100*s/ts:5.1f}% samples {100*i/ti:5.1f}% inst  {f}")
print(" smp%  inst%  lanes  line  source")
for d in sorted((d for d in data if d[0] == want), key=lambda d: d[1]):
    if 100 * d[4] / ti >= thr or 100 * d[3] / ts >= thr:
        print(f"{100*d[3]/ts:5.1f} {100*d[4]/ti:6.2f} {d[5]/d[4] if d[4] else 0:5.1f}  {d[1]:4d}  {d[2]}")
