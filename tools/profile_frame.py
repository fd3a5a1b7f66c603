"""Renders a few frames of one bench workload (for ncu captures): python tools/profile_frame.py headline 3"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_order_independent_transparency_b200 as oit  # noqa: E402
from bench import WORKLOADS  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "headline"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 3
W, H, kw, _ = WORKLOADS[name]
st = oit.State(**kw)
s = oit.Sample(st, W, H)
s.initScene()
ubo = oit.default_camera(W, H)
s.onRender(ubo)
s.synchronize()   # the first frame sizes the pair buffers (and is rendered again): keep it out of the capture window
for _ in range(frames):
    s.onRender(ubo)
    s.synchronize()
print(s.stats())
