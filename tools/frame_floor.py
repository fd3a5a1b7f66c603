"""Fixed cost of a frame: a tiny frame (few triangles, 256x160) rendered asynchronously many times -- what is left is the host's
enqueue cost and the device-side latency of the frame graph's nodes.  python tools/frame_floor.py [frames]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_order_independent_transparency_b200 as oit  # noqa: E402

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 400
for alg, aa in ((1, 4), (3, 0)):
    st = oit.State(algorithm=alg, aaType=aa, numObjects=8, subdiv=4)
    W, H = 256, 160
    s = oit.Sample(st, W, H)
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
    st_ = s.stats()
    print(f"alg {alg} aa {aa}: enqueue {1e6*(t1-t0)/frames:.1f} us/frame, total {1e6*(t2-t0)/frames:.1f} us/frame; launches {st_['kernelLaunches']}, "
          f"stages geo {st_['msGeometry']*1e3:.0f} us, clear {st_['msClear']*1e3:.0f}, color {st_['msColor']*1e3:.0f}, frame {st_['msFrame']*1e3:.0f}", flush=True)
    s.close()
