"""Small frames of several techniques for compute-sanitizer runs (memcheck / racecheck / initcheck):
compute-sanitizer --tool racecheck python tools/sanitize_frame.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_order_independent_transparency_b200 as oit  # noqa: E402

W, H = 112, 80
CASES = [dict(algorithm=1, aaType=4), dict(algorithm=3, aaType=0), dict(algorithm=4, aaType=2), dict(algorithm=6, aaType=1),
         dict(algorithm=2, aaType=3), dict(algorithm=5, aaType=1, percentTransparent=50), dict(algorithm=0, aaType=5),
         # linked list: the order-free kernel, and its pool-overflow path (tail blend through the shared-memory queue)
         dict(algorithm=1, aaType=0), dict(algorithm=1, aaType=0, linkedListAllocatedPerElement=1), dict(algorithm=1, aaType=1, linkedListAllocatedPerElement=2),
         dict(algorithm=1, aaType=5, percentTransparent=60)]
clip_only = len(sys.argv) > 1 and sys.argv[1] == "--clip-only"
for i, kw in enumerate(CASES * 2):
    inside = i >= len(CASES)   # second round: camera inside the cloud, so that triangles cross the near plane (oit_clip.cuh)
    if clip_only and not (inside and i % 3 == 0):
        continue
    st = oit.State(numObjects=24, subdiv=5, **kw)
    s = oit.Sample(st, W, H)
    s.initScene()
    ubo = oit.default_camera(W, H, eye=(0.3, -0.2, 0.6), center=(0.0, 0.1, -1.0), near=0.02) if inside else oit.default_camera(W, H)
    for _ in range(2):
        s.onRender(ubo)
    s.synchronize()
    print(kw, "inside" if inside else "outside", s.stats()["fragments"], flush=True)
    s.close()
print("done")
