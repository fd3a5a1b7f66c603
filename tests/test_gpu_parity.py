"""GPU parity tests: every call goes through the C ABI of liboit_b200.so (via the ctypes mirror) and is compared with
the CPU oracle on the same scene + UBO.  Bar: BIT-EXACT final image, colour samples, A-buffer contents and fragment
counts -- the CUDA path executes fragments in primitive order per pixel, which is the oracle's schedule, so even the
techniques that are racy in the reference (Simple, Spinlock, Loop64 tail) are deterministic here.  The only stated
tolerance is the linked list when its node pool overflows (which fragments overflow depends on allocation order)."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from helpers import make_oracle, max_channel_diff, scene_for, walk_lists  # noqa: E402

pytestmark = pytest.mark.gpu

ALGS = range(7)
AAS = range(6)
NCPU = os.cpu_count() or 1


def run_pair(oit, O, W, H, ubo=None, threads=1, fused_exact=True, **kw):
    st, verts, idx, ipo = scene_for(oit, **kw)
    ubo = ubo or oit.default_camera(W, H)
    # staged frame that keeps m_colorImage, so that the per-sample colour target can be compared as well
    s = oit.Sample(st, W, H, keepIntermediates=True)
    s.setScene(verts, idx, ipo)
    s.onRender(ubo)
    o, sd = make_oracle(O, st, W, H, verts, idx, ipo, ubo, threads)
    o.render(sd)
    # the default fused frame (colour pass + composite + resolve per tile, replayed as a CUDA graph) must give the same image
    f = oit.Sample(st, W, H)
    f.setScene(verts, idx, ipo)
    for _ in range(2):
        f.onRender(ubo)
        assert not fused_exact or np.array_equal(f.readColor(), o.final), "fused frame differs from the oracle"
    assert f.stats()["fragments"] == o.stats["fragments"]
    f.close()
    return s, o


def assert_frames_equal(s, o):
    gs, os_ = s.stats(), o.stats
    assert gs["fragments"] == os_["fragments"]
    assert gs["opaqueFragments"] == os_["opaqueFragments"]
    fin, ofin = s.readColor(), o.final
    assert np.array_equal(fin, ofin), f"{(fin != ofin).sum()} final pixels differ (max channel diff {max_channel_diff(fin, ofin)})"
    assert np.array_equal(s.colorSamples(), o.color_samples)
    assert gs["kernelLaunches"] > 0


# ---- the reference's own smoke matrix (test.py:34-51): 7 algorithms x {notail, tail} x 6 AA modes ---------------------
@pytest.mark.parametrize("tail", [0, 1])
@pytest.mark.parametrize("aa", AAS)
@pytest.mark.parametrize("alg", ALGS)
def test_matrix(oit_mod, oracle_mod, alg, aa, tail):
    s, o = run_pair(oit_mod, oracle_mod, 200, 128, algorithm=alg, aaType=aa, tailBlend=bool(tail), numObjects=160, subdiv=8)
    assert_frames_equal(s, o)
    gs, os_ = s.stats(), o.stats
    assert gs["fragmentsStored"] == os_["fragmentsStored"] and gs["fragmentsTail"] == os_["fragmentsTail"]
    s.close()


# ---- test.py's special sequences (test.py:13-30) at its 800x512 window ------------------------------------------------
SPECIAL = {
    "init": dict(),
    "interlock_unordered": dict(algorithm=5, interlockIsOrdered=False),
    "opaque": dict(percentTransparent=0),
    "opaque_msaa4": dict(percentTransparent=0, aaType=1),
    "opaque_ssaa4": dict(percentTransparent=0, aaType=2),
    "3objects": dict(numObjects=3),
    "lowsubdiv": dict(subdiv=2),
    "scaleMin": dict(scaleMin=1.0),
    "scaleWidth": dict(scaleWidth=10.0),
}


@pytest.mark.parametrize("name", sorted(SPECIAL))
def test_special_sequences(oit_mod, oracle_mod, name):
    s, o = run_pair(oit_mod, oracle_mod, 800, 512, threads=NCPU, **SPECIAL[name])
    assert_frames_equal(s, o)
    s.close()


@pytest.mark.parametrize("alg,aa,pct", [(1, 0, 50), (4, 1, 37), (5, 2, 80), (6, 4, 50), (2, 3, 99), (3, 5, 10), (0, 1, 1)])
def test_mixed_opaque_transparent(oit_mod, oracle_mod, alg, aa, pct):
    s, o = run_pair(oit_mod, oracle_mod, 240, 135, algorithm=alg, aaType=aa, percentTransparent=pct, numObjects=200, subdiv=6)
    assert_frames_equal(s, o)
    assert np.array_equal(s.download(oit_mod.BUF_DEPTH, np.float32).view(np.uint32), o.depth_samples.reshape(-1).view(np.uint32))
    s.close()


# ---- committed golden fixtures -----------------------------------------------------------------------------------------
def _golden():
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as G
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_small.npz"))
    return G, gold


def test_golden_fixtures(oit_mod):
    G, gold = _golden()
    names = sorted(k for k in gold.files if not k.endswith(("_F", "_abuf_sha256")))
    for name in names:
        a, aa, t, p = name.split("_")
        alg, aa, tail, pct = int(a[1:]), int(aa[2:]), int(t[1:]), int(p[1:])
        st, verts, idx, ipo = scene_for(oit_mod, algorithm=alg, aaType=aa, tailBlend=bool(tail), numObjects=G.OBJECTS, subdiv=G.SUBDIV, percentTransparent=pct)
        s = oit_mod.Sample(st, G.W, G.H)
        s.setScene(verts, idx, ipo)
        s.onRender(oit_mod.default_camera(G.W, G.H))
        assert s.stats()["fragments"] == int(gold[name + "_F"]), name
        assert np.array_equal(s.readColor(), gold[name]), name
        s.close()


# ---- A-buffer level: identical dumps ---------------------------------------------------------------------------------
@pytest.mark.parametrize("aa", [0, 1, 2, 4])
@pytest.mark.parametrize("alg", [0, 2, 3, 4, 5])
def test_kbuffer_dump_is_bit_exact(oit_mod, oracle_mod, alg, aa):
    s, o = run_pair(oit_mod, oracle_mod, 160, 120, algorithm=alg, aaType=aa, numObjects=200, subdiv=6, oitLayers=4)
    ab, oab = s.download(oit_mod.BUF_ABUFFER), o.abuffer
    if alg in (0, 4, 5):
        # slots beyond the per-pixel counter are never written (uninitialised in the reference too): compare what the composite reads
        cnt, ocnt = s.download(oit_mod.BUF_AUX), o.aux(0)
        assert np.array_equal(cnt, ocnt)
        stride = 4 if s.state.coverageShading() else 2
        P = s.bufW * s.bufH
        layers = s.msaa if s.sampleShading else 1
        a = ab.reshape(layers, 4, P, stride)
        b = oab.reshape(layers, 4, P, stride)
        valid = np.arange(4)[None, :, None] < np.minimum(cnt, 4).reshape(layers, 1, P)
        assert np.array_equal(a[valid], b[valid])
        if alg != 0:
            assert np.array_equal(s.download(oit_mod.BUF_AUXDEPTH), o.aux(2))
    elif alg == 2:
        P, layers = s.bufW * s.bufH, (s.msaa if s.sampleShading else 1)
        a, b = ab.reshape(layers, 2, 4, P), oab.reshape(layers, 2, 4, P)
        assert np.array_equal(a[:, 0], b[:, 0])                      # sorted depths
        valid = a[:, 0] != 0xFFFFFFFF
        assert np.array_equal(a[:, 1][valid], b[:, 1][valid])        # colours of the filled slots
    else:
        assert np.array_equal(ab, oab)
    s.close()


@pytest.mark.parametrize("aa", [0, 1, 2])
def test_linked_list_contents(oit_mod, oracle_mod, aa):
    s, o = run_pair(oit_mod, oracle_mod, 160, 120, algorithm=1, aaType=aa, numObjects=200, subdiv=6)
    assert s.download(oit_mod.BUF_COUNTER)[0] == o.aux(3)[0] == s.stats()["fragments"]
    got = walk_lists(s.download(oit_mod.BUF_ABUFFER), s.download(oit_mod.BUF_AUX))
    want = walk_lists(o.abuffer, o.aux(0))
    assert got == want  # same fragments, same per-pixel list order; only the node numbering differs
    s.close()


def test_weighted_targets(oit_mod, oracle_mod):
    for aa in (0, 1, 5):
        s, o = run_pair(oit_mod, oracle_mod, 160, 120, algorithm=6, aaType=aa, numObjects=200, subdiv=6)
        n = s.bufW * s.bufH * s.msaa
        assert np.array_equal(s.download(oit_mod.BUF_WACCUM, np.uint16), o.weighted(0))
        assert np.array_equal(s.download(oit_mod.BUF_WREVEAL, np.uint16)[:n], o.weighted(1))
        s.close()


# ---- composite on identical A-buffer dumps (north_star check #1), both directions --------------------------------------
@pytest.mark.parametrize("alg,aa,L", [(1, 0, 8), (1, 1, 2), (1, 2, 4), (0, 1, 8), (4, 0, 16), (5, 4, 8), (2, 0, 8), (3, 2, 8), (1, 0, 1), (4, 2, 32), (1, 4, 3)])
def test_composite_from_dump(oit_mod, oracle_mod, alg, aa, L):
    O = oracle_mod
    kw = dict(algorithm=alg, aaType=aa, oitLayers=L, numObjects=220, subdiv=6, linkedListAllocatedPerElement=12)
    W, H = 176, 112
    st, verts, idx, ipo = scene_for(oit_mod, **kw)
    ubo = oit_mod.default_camera(W, H)
    # (a) oracle colour pass -> dump -> CUDA composite + resolve
    o, sd = make_oracle(O, st, W, H, verts, idx, ipo, ubo)
    o.set_scene_data(sd)
    o.begin_frame(); o.draw_opaque(); o.draw_transparent()
    s = oit_mod.Sample(st, W, H)
    s.setScene(verts, idx, ipo)
    s.updateUniformBuffer(ubo)
    s.beginFrame()
    s.upload(oit_mod.BUF_ABUFFER, o.abuffer)
    s.upload(oit_mod.BUF_AUX, o.aux(0))
    s.upload(oit_mod.BUF_COLOR, o.color_samples)
    s.composite(); s.copyOffscreenToBackBuffer(); s.synchronize()
    o.composite(); o.resolve()
    assert np.array_equal(s.readColor(), o.final)
    # (b) CUDA colour pass -> dump -> oracle composite + resolve
    s.beginFrame(); s.drawOpaque(); s.drawTransparentColorOnly(); s.synchronize()
    o.begin_frame()
    o.abuffer[:] = s.download(oit_mod.BUF_ABUFFER)
    o.aux(0)[:] = s.download(oit_mod.BUF_AUX)
    o.color_samples[:] = s.colorSamples()
    o.composite(); o.resolve()
    s.composite(); s.copyOffscreenToBackBuffer(); s.synchronize()
    assert np.array_equal(s.readColor(), o.final)
    s.close()


# ---- edge cases ---------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("W,H", [(17, 9), (333, 197), (16, 16), (1, 1), (250, 3)])
@pytest.mark.parametrize("alg,aa", [(1, 1), (3, 0), (5, 2), (6, 3)])
def test_ragged_sizes(oit_mod, oracle_mod, W, H, alg, aa):
    s, o = run_pair(oit_mod, oracle_mod, W, H, algorithm=alg, aaType=aa, numObjects=64, subdiv=5)
    assert_frames_equal(s, o)
    s.close()


@pytest.mark.parametrize("L", [1, 2, 3, 16, 32])
@pytest.mark.parametrize("alg", [0, 1, 2, 3, 4, 5])
def test_layer_counts(oit_mod, oracle_mod, alg, L):
    s, o = run_pair(oit_mod, oracle_mod, 160, 100, algorithm=alg, oitLayers=L, numObjects=400, subdiv=5, aaType=1 if L <= 3 else 0)
    assert_frames_equal(s, o)
    s.close()


def test_empty_and_offscreen(oit_mod, oracle_mod):
    # camera looking away: nothing is rasterised, the frame is the clear colour (0.2 linear -> 124, alpha 51)
    ubo = oit_mod.default_camera(128, 80, eye=(0, 0, 12.0), center=(0, 0, 24.0))
    s, o = run_pair(oit_mod, oracle_mod, 128, 80, ubo=ubo, algorithm=1, numObjects=32, subdiv=4)
    assert s.stats()["fragments"] == 0 and np.all(s.readColor() == 0x337C7C7C)
    assert_frames_equal(s, o)
    s.close()
    # camera inside the cloud: triangles crossing the clip volume are rejected identically on both sides
    ubo = oit_mod.default_camera(128, 80, eye=(0, 0, 1.0), center=(0, 0, 0.0))
    s, o = run_pair(oit_mod, oracle_mod, 128, 80, ubo=ubo, algorithm=3, numObjects=300, subdiv=6)
    assert s.stats()["trianglesRejected"] > 0
    assert_frames_equal(s, o)
    s.close()


def test_linked_list_pool_overflow_tolerance(oit_mod, oracle_mod):
    """N=1: the pool holds W*H-1 fragments; which ones overflow is allocation-order dependent (racy in the reference,
    README.md:32).  Stated tolerance: the counters agree exactly and the images agree to a mean absolute channel
    difference below 12/255."""
    s, o = run_pair(oit_mod, oracle_mod, 160, 100, fused_exact=False, algorithm=1, linkedListAllocatedPerElement=1, numObjects=400, subdiv=6)
    gs, os_ = s.stats(), o.stats
    cap = 160 * 100
    assert gs["fragments"] == os_["fragments"] > cap
    assert gs["fragmentsStored"] == os_["fragmentsStored"] == cap - 1
    assert gs["fragmentsTail"] == os_["fragmentsTail"] == gs["fragments"] - (cap - 1)
    a = s.readColor().view(np.uint8).astype(np.int32)
    b = o.final.view(np.uint8).astype(np.int32)
    assert np.abs(a - b).mean() < 12.0
    s.close()


def test_reuse_and_idempotence(oit_mod, oracle_mod):
    st, verts, idx, ipo = scene_for(oit_mod, algorithm=4, aaType=1, numObjects=128, subdiv=6)
    s = oit_mod.Sample(st, 192, 128)
    s.setScene(verts, idx, ipo)
    ubo = oit_mod.default_camera(192, 128)
    s.onRender(ubo)
    a = s.readColor().copy()
    s.onRender(ubo)
    assert np.array_equal(a, s.readColor())
    ubo2 = oit_mod.default_camera(192, 128, eye=(3.0, 1.0, 11.0))
    s.onRender(ubo2)
    o, sd = make_oracle(oracle_mod, st, 192, 128, verts, idx, ipo, ubo2)
    o.render(sd)
    assert np.array_equal(s.readColor(), o.final)
    with pytest.raises(ValueError):
        s.clearTransparentLoop64()  # wrong algorithm, like the assert at oitRender.cpp:65
    s.close()


@pytest.mark.parametrize("mode", ["onchip_all", "no_onchip", "no_fuse", "no_graph", "sync_render", "no_pipeline", "layered_ll"])
@pytest.mark.parametrize("alg,aa,L", [(0, 0, 8), (0, 1, 4), (3, 0, 8), (3, 4, 2), (4, 1, 8), (5, 0, 8), (5, 4, 4), (6, 4, 8), (1, 1, 8)])
def test_frame_path_variants(oit_mod, oracle_mod, monkeypatch, mode, alg, aa, L):
    """Every way oit_render can execute a frame gives the oracle's image: k-buffer slice in shared memory for all
    techniques that support it, in HBM for all, staged (unfused) kernels, and plain stream launches instead of the graph."""
    monkeypatch.setenv({"onchip_all": "OIT_B200_ONCHIP_ALL", "no_onchip": "OIT_B200_NO_ONCHIP", "no_fuse": "OIT_B200_NO_FUSE",
                        "no_graph": "OIT_B200_NO_GRAPH", "sync_render": "OIT_B200_SYNC_RENDER", "no_pipeline": "OIT_B200_NO_PIPELINE",
                        "layered_ll": "OIT_B200_LAYERED_LL"}[mode], "1")
    s, o = run_pair(oit_mod, oracle_mod, 176, 120, algorithm=alg, aaType=aa, oitLayers=L, numObjects=180, subdiv=7)
    assert_frames_equal(s, o)
    s.close()


@pytest.mark.parametrize("alg,aa", [(1, 4), (3, 0), (6, 1)])
def test_frames_in_flight(oit_mod, alg, aa):
    """oit_render only enqueues: several frames with different cameras back to back (more than the UBO staging ring holds,
    starting on a fresh context whose pair buffers still have to grow), and what is read afterwards is the LAST camera's
    frame -- the same as rendering just that camera on its own context."""
    st = oit_mod.State(algorithm=alg, aaType=aa, numObjects=300, subdiv=8, linkedListAllocatedPerElement=40)
    W, H = 256, 160
    cams = [oit_mod.default_camera(W, H, eye=(0.3 * i, -0.2 * i, 12.0 - 0.5 * i)) for i in range(9)]
    s = oit_mod.Sample(st, W, H)
    s.initScene()
    for cam in cams:
        s.onRender(cam)
    got, got_stats = s.readColor(), s.stats()
    r = oit_mod.Sample(st, W, H)
    r.initScene()
    r.onRender(cams[-1])
    r.synchronize()
    assert np.array_equal(got, r.readColor())
    assert got_stats["fragments"] == r.stats()["fragments"]
    # the stage-by-stage entry points complete a frame that is still in flight before they start
    s.onRender(cams[0])
    s.updateUniformBuffer(cams[-1])
    s.beginFrame()
    s.drawOpaque()
    getattr(s, "drawTransparent" + {1: "LinkedList", 3: "Loop64", 6: "Weighted"}[alg])()
    s.copyOffscreenToBackBuffer()
    assert np.array_equal(s.readColor(), got)
    s.close()
    r.close()


@pytest.mark.parametrize("mode", ["graph", "no_graph"])
@pytest.mark.parametrize("alg,aa", [(1, 1), (4, 2), (2, 0)])
def test_pipelined_frames_alternate_buffer_sets(oit_mod, oracle_mod, monkeypatch, mode, alg, aa):
    """oit_render alternates between two sets of geometry buffers and runs a frame's vertex stage + binning on its own
    stream while the previous frame rasterises: EVERY frame of a sequence with changing cameras (some completed and read
    back, some left in flight) must be the oracle's frame of its own camera."""
    if mode == "no_graph":
        monkeypatch.setenv("OIT_B200_NO_GRAPH", "1")
    st, verts, idx, ipo = scene_for(oit_mod, algorithm=alg, aaType=aa, numObjects=150, subdiv=7)
    W, H = 208, 144
    cams = [oit_mod.default_camera(W, H, eye=(0.4 * i - 1.0, 0.15 * i, 11.0 + 0.3 * i)) for i in range(6)]
    s = oit_mod.Sample(st, W, H)
    s.setScene(verts, idx, ipo)
    for i, cam in enumerate(cams):
        s.onRender(cam)
        if i in (1, 3):
            continue   # left in flight: the next frame's geometry half overlaps this frame
        got, gs = s.readColor().copy(), s.stats()
        o, sd = make_oracle(oracle_mod, st, W, H, verts, idx, ipo, cam)
        o.render(sd)
        assert gs["fragments"] == o.stats["fragments"]
        assert np.array_equal(got, o.final), f"frame {i}: {(got != o.final).sum()} pixels differ"
        o.close()
    s.close()


def test_error_paths(oit_mod):
    s = oit_mod.Sample(oit_mod.State(algorithm=1), 64, 64)
    with pytest.raises(oit_mod.OitError) as e:
        s.onRender(oit_mod.default_camera(64, 64))
    assert e.value.code == -3  # OIT_ERR_NO_SCENE
    with pytest.raises(oit_mod.OitError):
        s.setScene(np.zeros((3, 10), np.float32), np.array([0, 1, 5], np.uint32), 3)  # index out of range
    with pytest.raises(oit_mod.OitError) as e:   # ... which leaves the context without a scene
        s.onRender(oit_mod.default_camera(64, 64))
    assert e.value.code == -3
    import torch
    bad = torch.tensor([0, 1, 7], dtype=torch.int32, device="cuda")
    dv = torch.zeros((3, 10), dtype=torch.float32, device="cuda")
    with pytest.raises(oit_mod.OitError):
        s.setSceneDevice(dv.data_ptr(), 3, bad.data_ptr(), 3, 3, keepalive=(dv, bad))   # same check for device-resident scenes
    s.setScene(np.zeros((3, 10), np.float32), np.array([0, 1, 2], np.uint32), 3)
    s.onRender(oit_mod.default_camera(64, 64))   # a valid scene afterwards renders
    with pytest.raises(oit_mod.OitError):
        s.upload(oit_mod.BUF_AUX, np.zeros(3, np.uint32))  # size mismatch
    s.close()
    import ctypes as C
    assert s.L.oit_buffer_size(None, 0, C.byref(C.c_size_t())) == -1


def test_abuffer_size_contract(oit_mod):
    # 499,875,840 bytes: Interlock, 16 layers, MSAA 4x pixel shading, 1920x1017 (reference screenshot; oit.cpp:155-156)
    s = oit_mod.Sample(oit_mod.State(algorithm=5, oitLayers=16, aaType=1), 1920, 1017)
    assert s.buffer_size(oit_mod.BUF_ABUFFER) == 499875840
    s.close()
    # README.md:29-37 bytes per pixel
    for alg, aa, L, per_pixel in [(0, 0, 8, 8 * 8), (0, 1, 8, 16 * 8), (2, 0, 8, 8 * 8), (3, 0, 4, 8 * 4), (4, 2, 8, 4 * 8 * 8), (1, 0, 8, 16 * 10)]:
        s = oit_mod.Sample(oit_mod.State(algorithm=alg, aaType=aa, oitLayers=L), 64, 48)
        assert s.buffer_size(oit_mod.BUF_ABUFFER) == 64 * 48 * per_pixel
        s.close()
