"""ctypes binding of the CPU oracle (oracle/oit_oracle.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboit_oracle.so")

OIT_SIMPLE, OIT_LINKEDLIST, OIT_LOOP, OIT_LOOP64, OIT_SPINLOCK, OIT_INTERLOCK, OIT_WEIGHTED = range(7)
AA_NONE, AA_MSAA_4X, AA_SSAA_4X, AA_SUPER_4X, AA_MSAA_8X, AA_SSAA_8X = range(6)


class OracleConfig(C.Structure):
    _fields_ = [
        ("algorithm", C.c_uint32),
        ("oitLayers", C.c_uint32),
        ("linkedListAllocatedPerElement", C.c_int32),
        ("percentTransparent", C.c_int32),
        ("tailBlend", C.c_uint32),
        ("interlockIsOrdered", C.c_uint32),
        ("numObjects", C.c_int32),
        ("subdiv", C.c_int32),
        ("scaleMin", C.c_float),
        ("scaleWidth", C.c_float),
        ("aaType", C.c_uint32),
        ("width", C.c_uint32),
        ("height", C.c_uint32),
    ]


class SceneData(C.Structure):
    _fields_ = [
        ("projViewMatrix", C.c_float * 16),
        ("viewMatrix", C.c_float * 16),
        ("viewMatrixInverseTranspose", C.c_float * 16),
        ("viewport", C.c_int32 * 3),
        ("linkedListAllocatedPerElement", C.c_uint32),
        ("alphaMin", C.c_float),
        ("alphaWidth", C.c_float),
        ("pad", C.c_float * 2),
    ]


class OracleStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("fragments", "fragmentsStored", "fragmentsTail", "opaqueFragments",
                                          "trianglesDrawn", "trianglesRejected", "llCounter")]


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "oit_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        L = _lib
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(OracleConfig)]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.oracle_set_scene.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        L.oracle_set_scene_data.argtypes = [C.c_void_p, C.POINTER(SceneData)]
        for f in ("oracle_begin_frame", "oracle_draw_opaque", "oracle_draw_transparent", "oracle_composite", "oracle_resolve"):
            getattr(L, f).argtypes = [C.c_void_p]
        L.oracle_render.argtypes = [C.c_void_p, C.POINTER(SceneData)]
        L.oracle_debug_invoke.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                          C.POINTER(C.c_float), C.c_float, C.c_float, C.POINTER(C.c_float)]
        for f, t in (("oracle_abuffer", C.c_uint32), ("oracle_color_samples", C.c_uint32), ("oracle_final", C.c_uint32),
                     ("oracle_depth_samples", C.c_float)):
            getattr(L, f).restype = C.POINTER(t)
            getattr(L, f).argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.oracle_aux.restype = C.POINTER(C.c_uint32)
        L.oracle_aux.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.oracle_weighted.restype = C.POINTER(C.c_uint16)
        L.oracle_weighted.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
        L.oracle_get_stats.argtypes = [C.c_void_p, C.POINTER(OracleStats)]
        L.oracle_buffer_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_uint32)] * 4
        L.oracle_scene_sizes.argtypes = [C.POINTER(OracleConfig)] + [C.POINTER(C.c_uint32)] * 3
        L.oracle_generate_scene.argtypes = [C.POINTER(OracleConfig), C.c_void_p, C.c_void_p]
        L.oracle_generate_scene_ex.argtypes = [C.POINTER(OracleConfig), C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_camera.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                    C.POINTER(C.c_float), C.c_float, C.c_float, C.POINTER(SceneData)]
        L.oracle_rand_canonical.restype = C.c_float
        L.oracle_rand_canonical.argtypes = [C.POINTER(C.c_uint64)]
        L.oracle_srgb_encode8.restype = C.c_uint32
        L.oracle_srgb_encode8.argtypes = [C.c_float]
        L.oracle_srgb_decode8.restype = C.c_float
        L.oracle_srgb_decode8.argtypes = [C.c_uint32]
        L.oracle_pack_color.restype = C.c_uint32
        L.oracle_pack_color.argtypes = [C.POINTER(C.c_float)]
        L.oracle_float_to_half.restype = C.c_uint16
        L.oracle_float_to_half.argtypes = [C.c_float]
        L.oracle_half_to_float.restype = C.c_float
        L.oracle_half_to_float.argtypes = [C.c_uint16]
        L.oracle_rop_blend.restype = C.c_uint32
        L.oracle_rop_blend.argtypes = [C.c_uint32, C.POINTER(C.c_float)]
    return _lib


def make_config(algorithm=OIT_SPINLOCK, oitLayers=8, linkedListAllocatedPerElement=10, percentTransparent=100,
                tailBlend=1, interlockIsOrdered=1, numObjects=1024, subdiv=16, scaleMin=0.1, scaleWidth=0.9,
                aaType=AA_NONE, width=1280, height=720):
    """Defaults = the reference's State defaults (oit.h:64-82)."""
    return OracleConfig(algorithm, oitLayers, linkedListAllocatedPerElement, percentTransparent, int(tailBlend),
                        int(interlockIsOrdered), numObjects, subdiv, scaleMin, scaleWidth, aaType, width, height)


STDLIB_LIBSTDCXX, STDLIB_MSVC = 0, 1


def generate_scene(cfg, stdlib=STDLIB_LIBSTDCXX):
    """initScene; `stdlib` picks whose std::default_random_engine drew the spheres (libstdc++: minstd_rand0, MSVC: mt19937)."""
    nv, ni, ipo = C.c_uint32(), C.c_uint32(), C.c_uint32()
    if lib().oracle_scene_sizes(C.byref(cfg), C.byref(nv), C.byref(ni), C.byref(ipo)) != 0:
        raise ValueError("bad scene parameters")
    verts = np.empty((nv.value, 10), np.float32)
    idx = np.empty(ni.value, np.uint32)
    lib().oracle_generate_scene_ex(C.byref(cfg), int(stdlib), verts.ctypes.data, idx.ctypes.data)
    return verts, idx, ipo.value


def camera(width, height, fov=45.0, eye=(0, 0, 12.0), center=(0, 0, 0), up=(0, 1, 0), near=0.1, far=100.0):
    sd = SceneData()
    f3 = C.c_float * 3
    lib().oracle_camera(width, height, fov, f3(*eye), f3(*center), f3(*up), near, far, C.byref(sd))
    return sd


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)


class Oracle:
    """Thin object wrapper: one oracle context."""

    def __init__(self, cfg, threads=1):
        self.cfg = cfg
        self.h = lib().oracle_create(C.byref(cfg))
        if not self.h:
            raise ValueError("oracle_create rejected the configuration")
        lib().oracle_set_threads(self.h, threads)
        w, h, m, ss = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().oracle_buffer_dims(self.h, C.byref(w), C.byref(h), C.byref(m), C.byref(ss))
        self.bufW, self.bufH, self.msaa, self.sampleShading = w.value, h.value, m.value, bool(ss.value)

    def close(self):
        if self.h:
            lib().oracle_destroy(self.h)
            self.h = None

    __del__ = close

    def set_scene(self, verts, idx, ipo):
        verts = np.ascontiguousarray(verts, np.float32)
        idx = np.ascontiguousarray(idx, np.uint32)
        r = lib().oracle_set_scene(self.h, verts.ctypes.data, verts.shape[0], idx.ctypes.data, idx.size, ipo)
        if r != 0:
            raise ValueError(f"oracle_set_scene failed ({r})")

    def set_scene_data(self, sd):
        lib().oracle_set_scene_data(self.h, C.byref(sd))

    def render(self, sd):
        lib().oracle_render(self.h, C.byref(sd))

    def begin_frame(self):
        lib().oracle_begin_frame(self.h)

    def draw_opaque(self):
        lib().oracle_draw_opaque(self.h)

    def draw_transparent(self):
        lib().oracle_draw_transparent(self.h)

    def composite(self):
        lib().oracle_composite(self.h)

    def resolve(self):
        lib().oracle_resolve(self.h)

    def debug_invoke(self, x, y, rgba, depth, pass_=1, sampleID=0, mask=1, viewZ=-10.0):
        out = (C.c_float * 4)()
        r = lib().oracle_debug_invoke(self.h, pass_, x, y, sampleID, mask, (C.c_float * 4)(*rgba), depth, viewZ, out)
        if r != 0:
            raise ValueError("oracle_debug_invoke failed")
        return np.array(out[:], np.float32)

    def _buf(self, fn, dtype, *a):
        n = C.c_size_t()
        p = fn(self.h, *a, C.byref(n))
        return _arr(p, n.value, dtype)

    @property
    def abuffer(self):
        return self._buf(lib().oracle_abuffer, np.uint32)

    def aux(self, which=0):
        return self._buf(lib().oracle_aux, np.uint32, which)

    @property
    def color_samples(self):
        return self._buf(lib().oracle_color_samples, np.uint32).reshape(self.bufH, self.bufW, self.msaa)

    @property
    def depth_samples(self):
        return self._buf(lib().oracle_depth_samples, np.float32).reshape(self.bufH, self.bufW, self.msaa)

    def weighted(self, which):
        return self._buf(lib().oracle_weighted, np.uint16, which)

    @property
    def final(self):
        return self._buf(lib().oracle_final, np.uint32).reshape(self.cfg.height, self.cfg.width)

    @property
    def stats(self):
        s = OracleStats()
        lib().oracle_get_stats(self.h, C.byref(s))
        return {n: getattr(s, n) for n, _ in OracleStats._fields_}


def bgra_to_rgba_image(final):
    """uint32 BGRA8 words (B in the low byte) -> HxWx4 uint8 RGBA array."""
    b = final.view(np.uint8).reshape(final.shape[0], final.shape[1], 4)
    return b[..., [2, 1, 0, 3]].copy()
