"""GPU tests at BASELINE.json's full sizes.  Configs 1-4 are still compared bit-for-bit with the (threaded) oracle; the 4K
headline case and the split-frame mode are additionally checked through size-independent properties: idempotence,
band-split invariance (any number of bands assembles to the single-GPU frame), and conservation of the fragment count."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import make_oracle, scene_for  # noqa: E402
from vk_order_independent_transparency_b200 import split_frame as SF  # noqa: E402

pytestmark = pytest.mark.gpu
NCPU = os.cpu_count() or 1

CONFIGS = {
    "cfg1_linkedlist_720p": (1280, 720, dict(algorithm=1, oitLayers=8, linkedListAllocatedPerElement=10)),
    "cfg2_loop64_1080p": (1920, 1080, dict(algorithm=3, oitLayers=8)),
    "cfg3_spinlock_msaa4_1080p": (1920, 1080, dict(algorithm=4, aaType=1)),
    "cfg3_spinlock_ssaa4_1080p": (1920, 1080, dict(algorithm=4, aaType=2)),
    "cfg3_interlock_msaa4_1080p": (1920, 1080, dict(algorithm=5, aaType=1)),
    "cfg3_interlock_ssaa4_1080p": (1920, 1080, dict(algorithm=5, aaType=2)),
    "cfg4_wboit_msaa8_4k": (3840, 2160, dict(algorithm=6, aaType=4)),
    # the configuration bench.py's headline number is quoted on (north-star target case)
    "headline_linkedlist_msaa8_4k": (3840, 2160, dict(algorithm=1, aaType=4)),
    # BASELINE config 5 (100 k spheres, Linked List, large A-buffer), as two cases the threaded oracle finishes in about a
    # minute, both sized so that the pool does NOT overflow (then the frame is defined and must match bit for bit):
    #   the first 20 k spheres of the scene at the full 3840x2160 (the generator is sequential: same spheres), N = 64
    #   ALL 100 k spheres (96 M triangles) at 960x540, N = 384
    "cfg5_first20k_spheres_4k_n64": (3840, 2160, dict(algorithm=1, numObjects=20000, linkedListAllocatedPerElement=64)),
    "cfg5_100k_spheres_960x540_n384": (960, 540, dict(algorithm=1, numObjects=100000, linkedListAllocatedPerElement=384)),
}


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_baseline_configs_match_oracle(oit_mod, oracle_mod, name):
    W, H, kw = CONFIGS[name]
    st, verts, idx, ipo = scene_for(oit_mod, **kw)
    ubo = oit_mod.default_camera(W, H)
    s = oit_mod.Sample(st, W, H)
    s.setScene(verts, idx, ipo)
    s.onRender(ubo)
    fin, gs = s.readColor(), s.stats()
    s.close()
    o, sd = make_oracle(oracle_mod, st, W, H, verts, idx, ipo, ubo, threads=NCPU)
    o.render(sd)
    assert gs["fragments"] == o.stats["fragments"] > 0
    assert gs["fragmentsStored"] == o.stats["fragmentsStored"] and gs["fragmentsTail"] == o.stats["fragmentsTail"]
    if name.startswith("cfg5") or name.startswith("headline"):
        assert gs["fragmentsTail"] == 0 and gs["llCounter"] == o.stats["llCounter"] == gs["fragments"]   # the pool did not overflow
    assert np.array_equal(fin, o.final), f"{(fin != o.final).sum()} pixels differ"
    o.close()


def test_config5_full_scene_overflow_regime(oit_mod):
    """BASELINE config 5 exactly as stated (100 k spheres, N = 128, 3840x2160): the pool (128 * W * H nodes) overflows, and
    WHICH fragments overflow depends on the allocation order (racy in the reference too, README.md:32), so the image is
    not defined bit for bit.  What is defined, and checked: every pool node is used exactly once, stored + tail-blended
    = all fragments, the counter = all fragments, the band split conserves the fragment count, and the frame is stable
    from run to run within the tolerance of test_linked_list_pool_overflow_tolerance (mean |diff| < 12 / 255)."""
    W, H = 3840, 2160
    st, verts, idx, ipo = scene_for(oit_mod, algorithm=1, numObjects=100000, linkedListAllocatedPerElement=128)
    ubo = oit_mod.default_camera(W, H)
    s = oit_mod.Sample(st, W, H)
    s.setScene(verts, idx, ipo)
    s.onRender(ubo)
    a, sa = s.readColor().copy(), s.stats()
    s.onRender(ubo)
    b, sb = s.readColor().copy(), s.stats()
    s.close()
    cap = 128 * W * H
    for st_ in (sa, sb):
        assert st_["fragments"] == sa["fragments"] > cap
        assert st_["fragmentsStored"] == cap - 1 and st_["fragmentsTail"] == st_["fragments"] - (cap - 1)
        assert st_["llCounter"] == st_["fragments"]
    d = np.abs(a.view(np.uint8).astype(np.int16) - b.view(np.uint8).astype(np.int16))
    assert d.mean() < 12.0
    # two bands: each has its own pool of 128 * W * localH nodes; the fragment count is conserved
    F = 0
    for band in range(2):
        sb2 = oit_mod.Sample(st, W, H, bandCount=2, bandIndex=band)
        sb2.setScene(verts, idx, ipo)
        sb2.onRender(ubo)
        F += sb2.stats()["fragments"]
        sb2.close()
    assert F == sa["fragments"]


def render_bands(oit, st, verts, idx, ipo, W, H, bands, strip=32):
    ubo = oit.default_camera(W, H)
    parts, F = [], 0
    for b in range(bands):
        s = oit.Sample(st, W, H, bandCount=bands, bandIndex=b, stripRows=strip)
        s.setScene(verts, idx, ipo)
        s.onRender(ubo)
        assert np.array_equal(s.globalRows(), SF.band_rows(H, bands, b, strip))
        assert s.localRowToGlobal(0) == SF.band_rows(H, bands, b, strip)[0]
        parts.append(s.readColor())
        F += s.stats()["fragments"]
        s.close()
    return SF.assemble(parts, H, W, strip), F


@pytest.mark.parametrize("alg,aa", [(1, 4), (3, 0), (5, 2), (6, 1), (2, 3)])
def test_band_split_invariance(oit_mod, alg, aa):
    W, H = 1280, 720
    st, verts, idx, ipo = scene_for(oit_mod, algorithm=alg, aaType=aa)
    full, F1 = render_bands(oit_mod, st, verts, idx, ipo, W, H, 1)
    for bands, strip in ((2, 32), (8, 32), (3, 64), (4, 16)):
        img, F = render_bands(oit_mod, st, verts, idx, ipo, W, H, bands, strip)
        assert F == F1
        assert np.array_equal(img, full), f"bands={bands}: {(img != full).sum()} pixels differ"


def test_headline_4k_msaa8_linked_list_properties(oit_mod):
    """default scene, Linked List, 3840x2160, 8x MSAA (the north-star target case)."""
    W, H = 3840, 2160
    st, verts, idx, ipo = scene_for(oit_mod, algorithm=1, aaType=4)
    ubo = oit_mod.default_camera(W, H)
    s = oit_mod.Sample(st, W, H)
    s.setScene(verts, idx, ipo)
    s.onRender(ubo)
    a, sa = s.readColor().copy(), s.stats()
    s.onRender(ubo)
    assert np.array_equal(a, s.readColor())                      # idempotent
    assert sa["fragments"] == s.stats()["fragments"] == sa["fragmentsStored"] == sa["llCounter"]  # nothing overflows at N=10
    assert sa["fragments"] > 20_000_000
    s.close()
    img, F = render_bands(oit_mod, st, verts, idx, ipo, W, H, 4)
    assert F == sa["fragments"] and np.array_equal(img, a)
    # background pixels keep the clear colour; covered pixels do not
    assert (a == 0x337C7C7C).mean() > 0.2 and (a != 0x337C7C7C).mean() > 0.3
