"""GPU parity on adversarial geometry (the sphere scene only has small, well-behaved triangles): random triangle soup under
an orthographic camera whose vertices sit on a half-pixel lattice, so that many edges pass exactly through sample
positions (top-left fill rule), plus degenerate triangles, slivers, triangles far larger than the frame (64-bit edge
functions, thousands of tiles each), vertices outside the guard band / depth range (rejected) and shared vertices."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import make_oracle  # noqa: E402

pytestmark = pytest.mark.gpu


def ortho_ubo(oit, W, H):
    """x, y in pixels map straight to the framebuffer (w = 1), z in [0, 1] is the depth; exact in float for power-of-two sizes."""
    sd = oit.SceneData()
    m = np.zeros((4, 4), np.float32)  # column major: m[c][r]
    m[0][0], m[3][0] = 2.0 / W, -1.0
    m[1][1], m[3][1] = 2.0 / H, -1.0
    m[2][2] = 1.0
    m[3][3] = 1.0
    sd.projViewMatrix[:] = list(m.reshape(-1))
    v = np.eye(4, dtype=np.float32)
    v[3][2] = -5.0  # view-space z only feeds the WBOIT weight
    sd.viewMatrix[:] = list(v.reshape(-1))
    sd.viewMatrixInverseTranspose[:] = list(np.eye(4, dtype=np.float32).reshape(-1))
    sd.alphaMin, sd.alphaWidth = 0.2, 0.3
    return sd


def soup(rng, W, H, n_small, n_big, n_special):
    tris = []
    for _ in range(n_small):  # small triangles on the half-pixel lattice
        c = rng.integers(0, [2 * W, 2 * H]) / 2.0
        tris.append(c + rng.integers(-24, 25, (3, 2)) / 2.0)
    for _ in range(n_big):  # far larger than the frame
        tris.append(rng.integers(-6000, 6000, (3, 2)).astype(np.float64) / 2.0)
    for _ in range(n_special):
        c = rng.integers(0, [W, H]).astype(np.float64) + 0.5  # pixel centres
        k = rng.integers(0, 5)
        if k == 0:    # degenerate: repeated vertex
            p = c + rng.integers(-9, 10, 2)
            tris.append(np.array([c, p, p]))
        elif k == 1:  # collinear
            d = rng.integers(-7, 8, 2).astype(np.float64)
            tris.append(np.array([c, c + d, c + 2 * d]))
        elif k == 2:  # axis-aligned right triangle with edges through pixel centres
            a, b = rng.integers(1, 12, 2)
            tris.append(np.array([c, c + [a, 0], c + [0, b]]))
        elif k == 3:  # sliver
            tris.append(np.array([c, c + [40, 0.5], c + [80, 0]]))
        else:         # the same triangle with the opposite winding
            a, b = rng.integers(1, 12, 2)
            tris.append(np.array([c, c + [0, b], c + [a, 0]]))
    tris = np.array(tris)
    n = len(tris)
    verts = np.zeros((n * 3, 10), np.float32)
    verts[:, 0:2] = tris.reshape(-1, 2)
    z = rng.random(n * 3) * 0.8 + 0.1
    bad = rng.random(n * 3) < 0.01  # a few vertices outside the depth clip volume -> their triangles are rejected
    z[bad] = rng.choice([-0.5, 1.5], bad.sum())
    verts[:, 2] = z
    verts[:, 3:6] = rng.standard_normal((n * 3, 3))
    verts[:, 6:10] = rng.random((n * 3, 4))
    idx = np.arange(n * 3, dtype=np.uint32)
    share = rng.integers(0, n * 3, n // 4)  # shared vertices
    idx[rng.integers(0, n * 3, n // 4)] = share
    return verts, idx


@pytest.mark.parametrize("alg,aa,pct", [(1, 0, 100), (1, 1, 100), (1, 4, 100), (5, 2, 100), (3, 0, 100), (6, 4, 100), (4, 5, 100), (0, 3, 100),
                                        (2, 1, 100), (1, 4, 60), (5, 0, 40)])
def test_triangle_soup_matches_oracle(oit_mod, oracle_mod, alg, aa, pct):
    W = H = 256
    rng = np.random.default_rng(1234 + alg * 10 + aa)
    verts, idx = soup(rng, W, H, 1500, 12, 400)
    st = oit_mod.State(algorithm=alg, aaType=aa, percentTransparent=pct, linkedListAllocatedPerElement=64, numObjects=len(idx) // 3)
    ubo = ortho_ubo(oit_mod, W * st.supersample, H * st.supersample)  # the UBO describes the (super-sampled) target
    if st.supersample == 2:
        verts = verts.copy()
        verts[:, 0:2] *= 2
    for keep in (True, False):  # staged frame with m_colorImage kept, and the default fused frame
        s = oit_mod.Sample(st, W, H, keepIntermediates=keep)
        s.setScene(verts, idx, 3)
        s.onRender(ubo)
        if keep:
            o, sd = make_oracle(oracle_mod, st, W, H, verts, idx, 3, ubo)
            o.render(sd)
            assert np.array_equal(s.colorSamples(), o.color_samples)
            assert o.stats["fragments"] > 50_000
        gs = s.stats()
        assert gs["fragments"] == o.stats["fragments"] and gs["opaqueFragments"] == o.stats["opaqueFragments"]
        assert gs["trianglesRejected"] > 0
        fin = s.readColor()
        assert np.array_equal(fin, o.final), f"{(fin != o.final).sum()} pixels differ (keepIntermediates={keep})"
        s.close()
