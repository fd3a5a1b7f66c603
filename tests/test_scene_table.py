"""the instanced scene input.  CPU: the 32-byte-per-object table is exactly what the flattened generator uses.
GPU: flattening the table on the device gives the frame (and statistics) of the host-flattened mesh, bit for bit."""
import numpy as np
import pytest

import vk_order_independent_transparency_b200 as oit


@pytest.mark.parametrize("kw", [dict(), dict(numObjects=37, subdiv=5, scaleMin=0.3, scaleWidth=0.2), dict(numObjects=1, subdiv=2)])
def test_table_matches_flattened_scene(kw):
    st = oit.State(**kw)
    verts, idx, ipo = oit.generate_scene(st)
    table = oit.generate_spheres(st)
    n = st.numObjects
    assert table.shape == (n, 8) and verts.shape[0] % n == 0
    v = verts.reshape(n, -1, 10)
    unit = v[0, :, 3:6]                                  # the normals are the unit sphere
    assert np.array_equal(v[:, :, 3:6], np.broadcast_to(unit, v[:, :, 3:6].shape))
    want = (unit[None, :, :] * table[:, None, 3:4]).astype(np.float32) + table[:, None, 0:3]   # separate multiply and add
    assert np.array_equal(v[:, :, 0:3], want.astype(np.float32))
    assert np.array_equal(v[:, :, 6:10], np.broadcast_to(table[:, None, 4:8], v[:, :, 6:10].shape))
    per = v.shape[1]
    assert np.array_equal(idx.reshape(n, ipo), idx[:ipo][None, :] + (np.arange(n, dtype=np.uint32) * per)[:, None])
    assert (table[:, 3] > 0).all() and (np.abs(table[:, 0:3]) <= 4.0).all()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(algorithm=1, aaType=4, linkedListAllocatedPerElement=64), dict(algorithm=3, numObjects=200, subdiv=6, percentTransparent=60),
                                dict(algorithm=6, aaType=1, numObjects=64, subdiv=9)])
def test_device_flattening_renders_the_same_frame(kw):
    st = oit.State(**kw)
    W, H = 320, 200
    ubo = oit.default_camera(W, H)
    a = oit.Sample(st, W, H)
    a.initScene()
    a.onRender(ubo)
    b = oit.Sample(st, W, H)
    b.setSceneSpheres(oit.generate_spheres(st))
    b.onRender(ubo)
    assert np.array_equal(a.readColor(), b.readColor())
    sa, sb = a.stats(), b.stats()
    for k in ("fragments", "fragmentsStored", "fragmentsTail", "opaqueFragments", "trianglesDrawn", "tilePairs"):
        assert sa[k] == sb[k], k
    # a different table of a different size on the same context, then back to the host path
    st2 = oit.State(**dict(kw, numObjects=23, subdiv=4))
    b.state = st2
    b.setSceneSpheres(oit.generate_spheres(st2), subdiv=4)
    b.onRender(ubo)
    c = oit.Sample(st2, W, H)
    c.initScene()
    c.onRender(ubo)
    assert np.array_equal(b.readColor(), c.readColor())
    verts, idx, ipo = oit.generate_scene(st)
    b.setScene(verts, idx, ipo)
    b.onRender(ubo)
    assert np.array_equal(a.readColor(), b.readColor())
    for s in (a, b, c):
        s.close()


@pytest.mark.gpu
def test_scene_spheres_errors():
    s = oit.Sample(oit.State(algorithm=1), 64, 64)
    with pytest.raises(oit.OitError):
        s.setSceneSpheres(np.zeros((0, 8), np.float32))
    with pytest.raises(oit.OitError):
        s.setSceneSpheres(np.zeros((4, 8), np.float32), subdiv=1)
    s.close()
