"""profiles/traffic.json entry of one workload from an `ncu --set full` capture of the frame kernel:

    python tools/make_traffic.py <capture.ncu-rep> <workload> [out.json]

Writes {workload: {dram_bytes_per_launch, dram_read, dram_write, l2 atomic / reduction sectors, kernel, capture,
kernel_source_sha16}} (merged into out.json, default profiles/traffic.json).  bench.py reads `dram_bytes_per_launch` for
`roofline.traffic` and REFUSES an entry whose `kernel_source_sha16` is not the hash of the kernel sources in the tree, so
run this on the same snapshot the capture was taken from (tools/final_profile.sh does, on the GPU box)."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    rep, workload = sys.argv[1], sys.argv[2]
    out = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "traffic.json")
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    # the longest captured launch of the frame kernel
    best = max((r for r in rows[2:] if "k_raster" in r[hdr.index("Kernel Name")]), key=lambda r: float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")))

    def val(name, scale_bytes=False):
        if name not in hdr:
            return None
        i = hdr.index(name)
        v = float(best[i].replace(",", ""))
        return v * UNIT.get(units[i], 1) if scale_bytes else v

    rd, wr = val("dram__bytes_read.sum", True), val("dram__bytes_write.sum", True)
    entry = {"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
             "lts_sectors_op_atom": val("lts__t_sectors_op_atom.sum"), "lts_sectors_op_red": val("lts__t_sectors_op_red.sum"),
             "kernel": best[hdr.index("Kernel Name")], "duration": best[hdr.index("gpu__time_duration.sum")] + " " + units[hdr.index("gpu__time_duration.sum")],
             "capture": os.path.basename(rep), "kernel_source_sha16": kernel_source_hash()}
    data = {}
    if os.path.exists(out):
        with open(out) as f:
            data = json.load(f)
    data[workload] = entry
    with open(out, "w") as f:
        json.dump(data, f, indent=1)
    print(json.dumps({workload: entry}))


if __name__ == "__main__":
    main()
