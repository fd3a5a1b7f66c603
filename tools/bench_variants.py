"""Times a few workloads with every library build under build/variants (developer tool for tuning constants)."""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = [("headline", dict(algorithm=1, aaType=4), 3840, 2160), ("ll_noaa_4k", dict(algorithm=1), 3840, 2160),
         ("loop64_1080p", dict(algorithm=3), 1920, 1080), ("spin_ssaa4_1080p", dict(algorithm=4, aaType=2), 1920, 1080),
         ("ll_720p", dict(algorithm=1), 1280, 720), ("wboit_msaa8_4k", dict(algorithm=6, aaType=4), 3840, 2160),
         ("simple_4k", dict(algorithm=0), 3840, 2160), ("loop32_4k", dict(algorithm=2), 3840, 2160), ("loop64_4k", dict(algorithm=3), 3840, 2160),
         ("spin_4k", dict(algorithm=4), 3840, 2160), ("interlock_4k", dict(algorithm=5), 3840, 2160), ("wboit_4k", dict(algorithm=6), 3840, 2160),
         ("interlock_msaa4_1080p", dict(algorithm=5, aaType=1), 1920, 1080)]
if os.environ.get("OIT_CASES"):
    CASES = [c for c in CASES if c[0] in os.environ["OIT_CASES"].split(",")]
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import vk_order_independent_transparency_b200 as oit
    out = {}
    for name, kw, W, H in CASES:
        st = oit.State(**kw)
        s = oit.Sample(st, W, H)
        s.initScene()
        ubo = oit.default_camera(W, H)
        for _ in range(3):
            s.onRender(ubo)
        ms, col = [], []
        for _ in range(8):
            s.onRender(ubo)
            t = s.stats()
            ms.append(t["msFrame"])
            col.append(t["msColor"])
        out[name] = (round(min(ms), 3), round(min(col), 3))
        s.close()
    print(json.dumps(out))
else:
    libs = sorted(glob.glob(os.path.join(ROOT, "build", "variants", "*.so"))) + [os.path.join(ROOT, "vk_order_independent_transparency_b200", "liboit_b200.so")]
    for lib in libs:
        env = dict(os.environ, OIT_B200_LIB=lib)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        print(os.path.basename(lib), r.stdout.strip() or r.stderr[-300:], flush=True)
