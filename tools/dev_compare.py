"""Developer sweep: CUDA path vs the CPU oracle over algorithms x AA modes x tail blend on a small frame.
Prints one line per case (mismatching pixels of the final image and of the colour samples, fragment counts)."""
import argparse
import itertools
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vk_order_independent_transparency_b200 as oit  # noqa: E402
from oracle import oracle_py as O  # noqa: E402


def run_case(alg, aa, tail, W, H, numObjects, subdiv, pct, layers, N, verbose=False):
    st = oit.State(algorithm=alg, aaType=aa, tailBlend=bool(tail), numObjects=numObjects, subdiv=subdiv,
                   percentTransparent=pct, oitLayers=layers, linkedListAllocatedPerElement=N)
    verts, idx, ipo = oit.generate_scene(st)
    ubo = oit.default_camera(W, H)
    s = oit.Sample(st, W, H, keepIntermediates=True)
    s.setScene(verts, idx, ipo)
    t0 = time.time()
    s.onRender(ubo)
    tg = time.time() - t0
    fin = s.readColor()
    col = s.colorSamples()
    gs = s.stats()

    ocfg = O.make_config(algorithm=alg, aaType=aa, tailBlend=tail, numObjects=numObjects, subdiv=subdiv,
                         percentTransparent=pct, oitLayers=layers, linkedListAllocatedPerElement=N, width=W, height=H)
    o = O.Oracle(ocfg)
    o.set_scene(verts, idx, ipo)
    sd = O.SceneData.from_buffer_copy(bytes(ubo))
    t0 = time.time()
    o.render(sd)
    to = time.time() - t0
    ofin, ocol, os_ = o.final, o.color_samples, o.stats
    dfin = int((fin != ofin).sum())
    dcol = int((col != ocol).sum())
    maxd = 0
    if dfin:
        a = fin.view(np.uint8).astype(np.int32)
        b = ofin.view(np.uint8).astype(np.int32)
        maxd = int(np.abs(a - b).max())
    ok = dfin == 0 and dcol == 0 and gs["fragments"] == os_["fragments"]
    print(f"{'OK  ' if ok else 'FAIL'} {oit.ALGORITHM_NAMES[alg]:10s} {oit.AA_NAMES[aa]:6s} tail={tail} "
          f"F gpu/ora={gs['fragments']}/{os_['fragments']} stored={gs['fragmentsStored']}/{os_['fragmentsStored']} "
          f"tailN={gs['fragmentsTail']}/{os_['fragmentsTail']} opaque={gs['opaqueFragments']}/{os_['opaqueFragments']} "
          f"finalDiff={dfin} (max {maxd}) sampleDiff={dcol}  gpu {gs['msFrame']:.3f} ms (wall {tg*1e3:.1f}) oracle {to*1e3:.0f} ms",
          flush=True)
    if verbose and not ok:
        ys, xs = np.nonzero(fin != ofin)
        for y, x in list(zip(ys, xs))[:8]:
            print(f"    ({x},{y}) gpu={fin[y, x]:08x} oracle={ofin[y, x]:08x}")
    s.close()
    o.close()
    return ok


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=320)
    ap.add_argument("--height", type=int, default=200)
    ap.add_argument("--objects", type=int, default=256)
    ap.add_argument("--subdiv", type=int, default=8)
    ap.add_argument("--pct", type=int, default=100)
    ap.add_argument("--layers", type=int, default=8)
    ap.add_argument("--N", type=int, default=10)
    ap.add_argument("--algs", default="0,1,2,3,4,5,6")
    ap.add_argument("--aas", default="0,1,2,3,4,5")
    ap.add_argument("--tails", default="1,0")
    ap.add_argument("-v", action="store_true")
    a = ap.parse_args()
    bad = 0
    for alg, aa, tail in itertools.product([int(x) for x in a.algs.split(",")], [int(x) for x in a.aas.split(",")],
                                           [int(x) for x in a.tails.split(",")]):
        try:
            bad += not run_case(alg, aa, tail, a.width, a.height, a.objects, a.subdiv, a.pct, a.layers, a.N, a.v)
        except Exception as e:  # keep sweeping
            bad += 1
            print(f"EXC  {oit.ALGORITHM_NAMES[alg]} {oit.AA_NAMES[aa]} tail={tail}: {e}", flush=True)
    print("failures:", bad)
    sys.exit(1 if bad else 0)
