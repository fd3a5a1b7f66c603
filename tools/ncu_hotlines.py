"""Top CUDA source lines of a kernel from an ncu report:  python tools/ncu_hotlines.py rep.ncu-rep k_raster [N]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, data, seen_kernel = None, None, [], None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        if seen_kernel is None:
            seen_kernel = r[1]
        fn = r[1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and len(r) > 10 and r[2] == "-" and fn == seen_kernel:
        try:
            data.append((cur, int(r[0]), r[1].strip()[:100], float(r[hdr["# Samples"]] or 0), float(r[hdr["Instructions Executed"]] or 0),
                         float(r[hdr["stall_barrier"]] or 0), float(r[hdr["stall_long_sb"]] or 0), float(r[hdr["stall_short_sb"]] or 0)))
        except ValueError:
            pass
ts, ti = sum(d[3] for d in data) or 1, sum(d[4] for d in data) or 1
print(seen_kernel, "samples", ts, "warp-inst", ti)
print(" smp%  inst%  bar%  lsb%  ssb%  file:line  source")
for d in sorted(data, key=lambda d: -d[3])[:top]:
    print(f"{100*d[3]/ts:5.1f} {100*d[4]/ti:6.1f} {100*d[5]/ts:5.1f} {100*d[6]/ts:5.1f} {100*d[7]/ts:5.1f}  {d[0]}:{d[1]}  {d[2]}")
