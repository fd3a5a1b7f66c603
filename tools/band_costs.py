"""Per-band colour-pass time and fragment count of a split frame, measured on ONE GPU (each band rendered on its own, no
exchange): shows how well the interleaved strips balance the bands.  python tools/band_costs.py [workload] [bands] [strips...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_order_independent_transparency_b200 as oit  # noqa: E402
from bench import WORKLOADS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "headline"
bands = int(sys.argv[2]) if len(sys.argv) > 2 else 8
strips = [int(x) for x in sys.argv[3:]] or [32]
W, H, kw, _ = WORKLOADS[name]
st = oit.State(**kw)
verts, idx, ipo = oit.generate_scene(st)
ubo = oit.default_camera(W, H)
for strip in strips:
    rows = []
    for b in range(bands):
        s = oit.Sample(st, W, H, bandCount=bands, bandIndex=b, stripRows=strip)
        s.setScene(verts, idx, ipo)
        col, geo = [], []
        for _ in range(6):
            s.onRender(ubo)
            t = s.stats()
            col.append(t["msColor"])
            geo.append(t["msGeometry"])
        rows.append((round(min(col), 4), round(min(geo), 4), t["fragments"], s.localRows))
        s.close()
    mx, mean = max(r[0] for r in rows), sum(r[0] for r in rows) / bands
    print(json.dumps({"workload": name, "bands": bands, "stripRows": strip, "color_ms_max": mx, "color_ms_mean": round(mean, 4),
                      "imbalance": round(mx / mean, 3), "per_band": rows}))
