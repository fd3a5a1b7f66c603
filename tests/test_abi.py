"""CPU tests of the drop-in boundary: liboit_b200.so loads, exports every symbol include/oit_b200.h declares, the PODs
have the documented sizes, the host-side harness (scene, camera) equals the oracle's, and -- on a box without a GPU --
every device entry point fails loudly instead of falling back to the CPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "oit_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(oit_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(oit_mod):
    lib = oit_mod.load_library()
    names = header_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/oit_b200.h but not exported"
    assert sorted(oit_mod.ABI_SYMBOLS) == names
    assert lib.oit_abi_version() == 1


def test_srgb_encoder_is_exact_for_every_float(oit_mod):
    """The frame kernels encode linear -> sRGB8 with a bucket table + one threshold compare (oit_device.cuh: enc8) instead of
    the GLSL's pow(); the library proves on the host that this equals the threshold definition for ALL floats in [0, 1]."""
    lib = oit_mod.load_library()
    lib.oit_selfcheck_srgb_encoder.argtypes = [C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.oit_selfcheck_srgb_encoder.restype = C.c_int
    n, bad = C.c_uint64(), C.c_uint64()
    assert lib.oit_selfcheck_srgb_encoder(1, C.byref(n), C.byref(bad)) == 0
    assert n.value == 0x3F800000 + 1 + 11 and bad.value == 0


def test_pod_layouts(oit_mod):
    assert C.sizeof(oit_mod.SceneData) == 224          # shaders/common.h:77-92 (std140)
    assert oit_mod.SceneData.viewport.offset == 192 and oit_mod.SceneData.linkedListAllocatedPerElement.offset == 204
    assert oit_mod.SceneData.alphaMin.offset == 208
    assert C.sizeof(oit_mod.OitConfig) == 21 * 4
    cfg = oit_mod.OitConfig()
    oit_mod.load_library().oit_default_config(C.byref(cfg))
    # State{} defaults (oit.h:64-82)
    assert (cfg.algorithm, cfg.oitLayers, cfg.linkedListAllocatedPerElement, cfg.percentTransparent) == (4, 8, 10, 100)
    assert (cfg.tailBlend, cfg.interlockIsOrdered, cfg.numObjects, cfg.subdiv, cfg.aaType) == (1, 1, 1024, 16, 0)
    assert abs(cfg.scaleMin - 0.1) < 1e-7 and abs(cfg.scaleWidth - 0.9) < 1e-7 and cfg.bandCount == 1


def test_state_antialiasing_table(oit_mod):
    # oit.h:84-115
    want = {0: (1, False, 1), 1: (4, False, 1), 2: (4, True, 1), 3: (1, False, 2), 4: (8, False, 1), 5: (8, True, 1)}
    for aa, (m, ss, sup) in want.items():
        st = oit_mod.State(aaType=aa)
        assert (st.msaa, st.sampleShading, st.supersample) == (m, ss, sup)
        assert st.coverageShading() == (m > 1 and not ss)
    with pytest.raises(ValueError):
        oit_mod.State(aaType=9)


@pytest.mark.parametrize("kw", [dict(numObjects=7, subdiv=2), dict(numObjects=3, subdiv=16, scaleMin=1.0), dict(numObjects=5, subdiv=5, scaleWidth=10.0)])
def test_scene_generator_matches_oracle(oit_mod, oracle_mod, kw):
    v, i, ipo = oit_mod.generate_scene(oit_mod.State(**kw))
    ov, oi, oipo = oracle_mod.generate_scene(oracle_mod.make_config(**kw))
    assert ipo == oipo and np.array_equal(i, oi) and np.array_equal(v.view(np.uint32), ov.view(np.uint32))


def test_camera_matches_oracle(oit_mod, oracle_mod):
    for (w, h, near, far) in [(1280, 720, 0.1, 100.0), (800, 512, 0.001, 1e8), (3840, 2160, 0.1, 100.0)]:
        assert bytes(oit_mod.default_camera(w, h, near=near, far=far)) == bytes(oracle_mod.camera(w, h, near=near, far=far))
    sd = oit_mod.default_camera(1280, 720)
    assert sd.viewport[2] == 1280 * 720 and abs(sd.alphaMin - 0.2) < 1e-7 and abs(sd.alphaWidth - 0.3) < 1e-7


def test_bad_arguments_are_reported_not_fatal(oit_mod):
    lib = oit_mod.load_library()
    h = C.c_void_p()
    for bad in (dict(algorithm=7), dict(aaType=6), dict(oitLayers=0), dict(oitLayers=33)):
        st = oit_mod.State()
        cfg = st.to_config(64, 64)
        for k, v in bad.items():
            setattr(cfg, k, v)
        assert lib.oit_create(C.byref(cfg), C.byref(h)) == -1
        assert lib.oit_last_error(None)
    cfg = oit_mod.State().to_config(0, 64)
    assert lib.oit_create(C.byref(cfg), C.byref(h)) == -1
    assert lib.oit_destroy(None) == 0


def test_no_cpu_fallback_without_gpu(oit_mod):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(oit_mod.OitError) as e:
        oit_mod.Sample(oit_mod.State(), 64, 64)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vk_order_independent_transparency_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read().lower().replace("oracle's", ""), f


def test_save_png_roundtrip(oit_mod, tmp_path):
    """save_png (the sequencer's saveImageToFile, main.cpp:887-891) writes the frame's bytes unchanged: BGRA -> RGBA PNG."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(7)
    frame = rng.integers(0, 2**32, size=(37, 53), dtype=np.uint64).astype(np.uint32)
    path = tmp_path / "frame.png"
    oit_mod.save_png(frame, str(path))
    back = np.asarray(Image.open(path))
    assert back.shape == (37, 53, 4) and back.dtype == np.uint8
    assert np.array_equal(back, oit_mod.bgra_to_rgba_image(frame))
