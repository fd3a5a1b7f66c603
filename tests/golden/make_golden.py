"""Generates tests/golden/oracle_small.npz from the CPU oracle.

The reference has no golden images (test.py only checks the exit code) and cannot run here (no Vulkan), so these
fixtures pin the ORACLE: CPU tests check that the oracle still reproduces them, GPU tests check the CUDA path against
them.  Run from the repo root:  python tests/golden/make_golden.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle_py as O  # noqa: E402

W, H, OBJECTS, SUBDIV = 160, 96, 96, 6
CASES = [(alg, aa, 1) for alg in range(7) for aa in (0, 1, 2)] + [(1, 3, 1), (3, 4, 1), (4, 5, 1), (5, 4, 0), (0, 0, 0), (6, 3, 1)]


def case_name(alg, aa, tail, pct=100):
    return f"a{alg}_aa{aa}_t{tail}_p{pct}"


def render(alg, aa, tail, pct=100):
    cfg = O.make_config(algorithm=alg, aaType=aa, tailBlend=tail, numObjects=OBJECTS, subdiv=SUBDIV, percentTransparent=pct, width=W, height=H)
    verts, idx, ipo = O.generate_scene(cfg)
    o = O.Oracle(cfg)
    o.set_scene(verts, idx, ipo)
    o.render(O.camera(W, H))
    fin = o.final.copy()
    digest = hashlib.sha256(o.abuffer.tobytes() if alg not in (1, 6) else b"").hexdigest()
    F = o.stats["fragments"]
    o.close()
    return fin, digest, F


if __name__ == "__main__":
    out = {}
    for alg, aa, tail in CASES:
        fin, digest, F = render(alg, aa, tail)
        n = case_name(alg, aa, tail)
        out[n] = fin
        out[n + "_abuf_sha256"] = np.frombuffer(bytes.fromhex(digest), np.uint8)
        out[n + "_F"] = np.uint64(F)
    for alg in (1, 5):
        fin, digest, F = render(alg, 1, 1, 60)
        n = case_name(alg, 1, 1, 60)
        out[n], out[n + "_abuf_sha256"], out[n + "_F"] = fin, np.frombuffer(bytes.fromhex(digest), np.uint8), np.uint64(F)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_small.npz"), **out)
    print("wrote", len(out) // 3, "cases")
