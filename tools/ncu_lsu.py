"""Shared-memory wavefronts (L1 data pipe) and global L1 tag requests per source REGION and per line of one kernel:
   python tools/ncu_lsu.py rep.ncu-rep k_raster_ll regions.txt [top_lines]
regions.txt as for tools/ncu_regions.py.  The frame kernels are bound by the LSU data pipe as much as by issue slots."""
import csv
import subprocess
import sys
from collections import defaultdict

rep, kern, regf = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
regions = []
for ln in open(regf):
    f = ln.split()
    if len(f) == 4 and not ln.startswith("#"):
        regions.append((f[0], f[1], int(f[2]), int(f[3])))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr, seen, fn = None, None, None, None
acc = defaultdict(lambda: [0.0] * 4)
lines = defaultdict(lambda: [0.0] * 4 + [""])
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
        def g(name):
            try:
                return float(r[hdr[name]] or 0)
            except (ValueError, KeyError):
                return 0.0
        try:
            line = int(r[0])
        except ValueError:
            continue
        v = [g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Excessive"), g("L1 Tag Requests Global"), g("Instructions Executed")]
        name = next((n for n, f, a, b in regions if f == cur and a <= line <= b), f"other:{cur}")
        for k in range(4):
            acc[name][k] += v[k]
            lines[(cur, line)][k] += v[k]
        lines[(cur, line)][4] = r[1].strip()[:90]
tot = [sum(v[k] for v in acc.values()) or 1 for k in range(4)]
print(f"{seen}: shared wavefronts {tot[0]:.0f} (excessive {tot[1]:.0f}), global tag requests {tot[2]:.0f}, warp-inst {tot[3]:.0f}")
print(f"{'region':<34} {'smemWF%':>8} {'excess%':>8} {'gtag%':>7} {'inst%':>6}")
for n, v in sorted(acc.items(), key=lambda kv: -(kv[1][0] + kv[1][2])):
    print(f"{n:<34} {100*v[0]/tot[0]:8.1f} {100*v[1]/tot[0]:8.1f} {100*v[2]/tot[2]:7.1f} {100*v[3]/tot[3]:6.1f}")
print("\nhottest lines by shared wavefronts + global tag requests")
for (f, l), v in sorted(lines.items(), key=lambda kv: -(kv[1][0] + kv[1][2]))[:top]:
    print(f"{100*v[0]/tot[0]:6.1f} {100*v[1]/tot[0]:6.1f} {100*v[2]/tot[2]:6.1f}  {f}:{l}  {v[4]}")
