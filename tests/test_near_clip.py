"""Near-plane clipping ("standard 0 <= z_clip <= w_clip near/far clip").  The default camera never
reaches the near plane, so these scenes put the camera INSIDE the geometry.

CPU: the oracle, looking out from inside one coarse sphere whose triangles are so large that several have vertices behind
the camera while they cover part of the screen -- every view ray leaves a closed surface exactly once, so every pixel must
receive exactly ONE fragment: a missing clip leaves holes, a clip that is not watertight leaves holes or double hits.
GPU: the CUDA path against the oracle, bit for bit, on such scenes and inside the default sphere cloud."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import vk_order_independent_transparency_b200 as oit  # noqa: E402
from helpers import make_oracle  # noqa: E402

W, H = 160, 112


def inside_sphere_scene(subdiv, look=(0.3, 0.2, -1.0), offset=(0.02, -0.03, 0.01), **kw):
    st = oit.State(numObjects=1, subdiv=subdiv, scaleMin=1.0, scaleWidth=0.0, **kw)
    verts, idx, ipo = oit.generate_scene(st)
    c = oit.generate_spheres(st)[0]
    eye = tuple(float(c[k] + offset[k] * c[3]) for k in range(3))
    center = tuple(eye[k] + look[k] for k in range(3))
    ubo = oit.default_camera(W, H, fov=70.0, eye=eye, center=center, near=0.01 * float(c[3]), far=100.0)
    return st, verts, idx, ipo, ubo


@pytest.mark.parametrize("subdiv", [2, 3, 5])
@pytest.mark.parametrize("look", [(0.3, 0.2, -1.0), (1.0, 0.1, 0.2), (-0.2, -1.0, 0.4)])
def test_oracle_inside_a_closed_surface_every_pixel_once(subdiv, look):
    from oracle import oracle_py as O
    st, verts, idx, ipo, ubo = inside_sphere_scene(subdiv, look, algorithm=oit.OIT_LINKEDLIST, linkedListAllocatedPerElement=4)
    o, sd = make_oracle(O, st, W, H, verts, idx, ipo, ubo, 1)
    o.render(sd)
    assert o.stats["fragments"] == W * H, (o.stats["fragments"], W * H)
    heads = o.aux(0).reshape(-1)[: W * H]
    assert (heads != 0).all()                                   # every pixel has a list ...
    nodes = o.abuffer.reshape(-1, 4)
    assert (nodes[heads, 3] == 0).all()                         # ... of exactly one node
    o.close()


def test_oracle_band_threads_agree_on_clipped_scenes():
    """The oracle's band-parallel mode (what cpu_baseline times) clips per thread: same image as the sequential walk."""
    from oracle import oracle_py as O
    st = oit.State(algorithm=oit.OIT_SPINLOCK, aaType=oit.AA_MSAA_4X, numObjects=200, subdiv=6)
    verts, idx, ipo = oit.generate_scene(st)
    ubo = oit.default_camera(W, H, eye=(0.2, 0.1, 0.8), center=(0.0, 0.0, -1.0), near=0.05)
    imgs, frags = [], []
    for threads in (1, 5):
        o, sd = make_oracle(O, st, W, H, verts, idx, ipo, ubo, threads)
        o.render(sd)
        imgs.append(o.final.copy())
        frags.append((o.stats["fragments"], o.stats["trianglesRejected"]))
        o.close()
    assert np.array_equal(imgs[0], imgs[1]) and frags[0] == frags[1] and frags[0][0] > 0


def test_oracle_clipped_scene_differs_from_rejecting(monkeypatch):
    """Sanity of the test itself: the scene really has triangles with vertices behind the near plane."""
    st, verts, idx, ipo, ubo = inside_sphere_scene(2)
    M = np.array(list(ubo.projViewMatrix), np.float32).reshape(4, 4)    # column-major: M[c][r]
    zc = verts[:, 0:1] * M[0, 2] + verts[:, 1:2] * M[1, 2] + verts[:, 2:3] * M[2, 2] + M[3, 2]
    behind = (zc[:, 0] < 0)[idx.reshape(-1, 3)]
    crossing = behind.any(axis=1) & ~behind.all(axis=1)
    assert crossing.sum() >= 2


GPU_CASES = [dict(algorithm=1, aaType=0, linkedListAllocatedPerElement=4), dict(algorithm=1, aaType=4, linkedListAllocatedPerElement=8),
             dict(algorithm=5, aaType=1), dict(algorithm=3, aaType=2), dict(algorithm=6, aaType=5), dict(algorithm=2, aaType=3),
             dict(algorithm=4, aaType=0, percentTransparent=0)]


@pytest.mark.gpu
@pytest.mark.parametrize("kw", GPU_CASES)
@pytest.mark.parametrize("subdiv", [2, 4])
def test_cuda_matches_oracle_inside_a_sphere(kw, subdiv):
    from oracle import oracle_py as O
    st, verts, idx, ipo, ubo = inside_sphere_scene(subdiv, **kw)
    o, sd = make_oracle(O, st, W, H, verts, idx, ipo, ubo, 1)
    o.render(sd)
    for keep in (True, False):                       # staged kernels and the fused frame
        s = oit.Sample(st, W, H, keepIntermediates=keep)
        s.setScene(verts, idx, ipo)
        s.onRender(ubo)
        assert np.array_equal(s.readColor(), o.final), f"{(s.readColor() != o.final).sum()} pixels differ (keepIntermediates={keep})"
        gs = s.stats()
        assert gs["fragments"] == o.stats["fragments"] and gs["trianglesRejected"] == o.stats["trianglesRejected"]
        s.close()
    o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(algorithm=1, aaType=4, linkedListAllocatedPerElement=40), dict(algorithm=4, aaType=1, percentTransparent=60),
                                dict(algorithm=6, aaType=0), dict(algorithm=3, aaType=0)])
def test_cuda_matches_oracle_inside_the_sphere_cloud(kw):
    """The default scene with the camera in the middle of the cloud: hundreds of triangles cross the near plane."""
    from oracle import oracle_py as O
    st = oit.State(numObjects=400, subdiv=8, **kw)
    verts, idx, ipo = oit.generate_scene(st)
    w, h = 320, 200
    ubo = oit.default_camera(w, h, eye=(0.4, -0.3, 1.0), center=(0.0, 0.2, -1.0), near=0.05)
    o, sd = make_oracle(O, st, w, h, verts, idx, ipo, ubo, os.cpu_count() or 1)
    o.render(sd)
    s = oit.Sample(st, w, h)
    s.setScene(verts, idx, ipo)
    s.onRender(ubo)
    assert np.array_equal(s.readColor(), o.final), f"{(s.readColor() != o.final).sum()} pixels differ"
    assert s.stats()["fragments"] == o.stats["fragments"]
    # split frame: the pieces are binned per band like any other triangle
    parts = []
    import vk_order_independent_transparency_b200.split_frame as SF
    for b in range(3):
        t = oit.Sample(st, w, h, bandCount=3, bandIndex=b, stripRows=16)
        t.setScene(verts, idx, ipo)
        t.onRender(ubo)
        parts.append(t.readColor())
        t.close()
    assert np.array_equal(SF.assemble(parts, h, w, 16), o.final)
    s.close()
    o.close()


@pytest.mark.gpu
def test_clip_table_grows():
    """More pieces than the clip table's initial 4096 entries: the first attempt raises the overflow flag, the table grows
    and the frame is rendered again (like the pair buffers) -- the result is still the oracle's, bit for bit."""
    from oracle import oracle_py as O
    st = oit.State(algorithm=oit.OIT_LOOP64, numObjects=1024, subdiv=16, scaleMin=1.0, scaleWidth=0.0)
    verts, idx, ipo = oit.generate_scene(st)
    w, h = 320, 200
    ubo = oit.default_camera(w, h, eye=(0.0, 0.0, 0.5), center=(0.0, 0.0, -1.0), near=0.05)
    M = np.array(list(ubo.projViewMatrix), np.float32).reshape(4, 4)
    zc = verts[:, 0:1] * M[0, 2] + verts[:, 1:2] * M[1, 2] + verts[:, 2:3] * M[2, 2] + M[3, 2]
    behind = (zc[:, 0] < 0)[idx.reshape(-1, 3)]
    assert (behind.any(axis=1) & ~behind.all(axis=1)).sum() > 4096          # the scene really needs more entries
    o, sd = make_oracle(O, st, w, h, verts, idx, ipo, ubo, os.cpu_count() or 1)
    o.render(sd)
    s = oit.Sample(st, w, h)
    s.setScene(verts, idx, ipo)
    for _ in range(2):
        s.onRender(ubo)
        assert np.array_equal(s.readColor(), o.final), f"{(s.readColor() != o.final).sum()} pixels differ"
    gs = s.stats()
    assert gs["fragments"] == o.stats["fragments"] and gs["trianglesRejected"] == o.stats["trianglesRejected"]
    s.close()
    o.close()
