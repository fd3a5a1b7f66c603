"""One band of a split frame rendered asynchronously on ONE GPU without the exchange: frame time of a stream of frames
against the colour-pass time -- what the frame pipeline itself (geometry half on its own stream, graph nodes) costs at the
small per-band frame of an 8-GPU split.  python tools/band_async.py [workload] [bands] [band] [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_order_independent_transparency_b200 as oit  # noqa: E402
from bench import WORKLOADS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "headline"
bands = int(sys.argv[2]) if len(sys.argv) > 2 else 8
band = int(sys.argv[3]) if len(sys.argv) > 3 else 0
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 200
W, H, kw, _ = WORKLOADS[name]
st = oit.State(**kw)
s = oit.Sample(st, W, H, bandCount=bands, bandIndex=band)
s.initScene()
ubo = oit.default_camera(W, H)
for _ in range(10):
    s.onRender(ubo)
s.synchronize()
t0 = time.perf_counter()
for _ in range(frames):
    s.onRender(ubo)
t1 = time.perf_counter()
s.synchronize()
t2 = time.perf_counter()
col, geo = [], []
for _ in range(5):
    s.onRender(ubo)
    t = s.stats()
    col.append(t["msColor"])
    geo.append(t["msGeometry"])
print(f"{name} band {band}/{bands}: {1e3*(t2-t0)/frames:.4f} ms/frame in a stream of frames (host enqueue {1e3*(t1-t0)/frames:.4f}); "
      f"one at a time: colour {min(col):.4f}, geometry {min(geo):.4f}, launches {t['kernelLaunches']}", flush=True)
s.close()
