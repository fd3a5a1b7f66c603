"""B200-native order-independent-transparency hot path (drop-in for oitRender.cpp's clear / draw / composite stages).

This package is a thin host-side mirror of the reference's `Sample` interface (oit.h:379-425) over the C ABI of
`liboit_b200.so` (include/oit_b200.h).  All work happens in hand-written sm_100a CUDA kernels; there is NO CPU
fallback: importing works without a GPU (so the ABI can be inspected), but creating a renderer without the library or
without a CUDA device raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboit_b200.so")

# shaders/common.h:44-63
OIT_SIMPLE, OIT_LINKEDLIST, OIT_LOOP, OIT_LOOP64, OIT_SPINLOCK, OIT_INTERLOCK, OIT_WEIGHTED = range(7)
AA_NONE, AA_MSAA_4X, AA_SSAA_4X, AA_SUPER_4X, AA_MSAA_8X, AA_SSAA_8X = range(6)
ALGORITHM_NAMES = ["simple", "linkedlist", "loop", "loop64", "spinlock", "interlock", "weighted"]  # test.py:34-42
AA_NAMES = ["noaa", "msaa4", "ssaa4", "super4", "msaa8", "ssaa8"]  # test.py:44

(BUF_ABUFFER, BUF_AUX, BUF_AUXSPIN, BUF_AUXDEPTH, BUF_COUNTER, BUF_COLOR, BUF_DEPTH, BUF_WACCUM, BUF_WREVEAL,
 BUF_FINAL) = range(10)

# every symbol include/oit_b200.h declares
ABI_SYMBOLS = [
    "oit_abi_version", "oit_default_config", "oit_create", "oit_destroy", "oit_last_error", "oit_get_config",
    "oit_get_dims", "oit_set_scene", "oit_set_scene_device", "oit_scene_sizes", "oit_generate_scene",
    "oit_default_camera", "oit_render", "oit_set_scene_data", "oit_begin_frame", "oit_draw_opaque",
    "oit_draw_transparent", "oit_composite", "oit_resolve", "oit_synchronize", "oit_buffer_size", "oit_download",
    "oit_upload", "oit_device_ptr", "oit_read_color", "oit_get_stats", "oit_stream", "oit_local_row_to_global",
    "oit_band_gather_unique_id", "oit_enable_band_gather", "oit_band_peer_export", "oit_band_peer_enable",
    "oit_band_peer_disable", "oit_generate_spheres", "oit_set_scene_spheres", "oit_selfcheck_srgb_encoder",
]
BUF_FRAME = 10
STDLIB_LIBSTDCXX, STDLIB_MSVC = 0, 1  # OIT_CFG_SCENE_STDLIB: minstd_rand0 (libstdc++) or mt19937 (MSVC) scene RNG


class OitError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"liboit_b200 error {code}: {msg}")
        self.code = code


class OitConfig(C.Structure):
    _fields_ = [
        ("algorithm", C.c_uint32), ("oitLayers", C.c_uint32), ("linkedListAllocatedPerElement", C.c_int32),
        ("percentTransparent", C.c_int32), ("tailBlend", C.c_uint32), ("interlockIsOrdered", C.c_uint32),
        ("numObjects", C.c_int32), ("subdiv", C.c_int32), ("scaleMin", C.c_float), ("scaleWidth", C.c_float),
        ("aaType", C.c_uint32), ("width", C.c_uint32), ("height", C.c_uint32), ("device", C.c_int32),
        ("bandCount", C.c_uint32), ("bandIndex", C.c_uint32), ("stripRows", C.c_uint32), ("reserved", C.c_uint32 * 4),
    ]


class SceneData(C.Structure):
    """shaderio::SceneData (shaders/common.h:77-92), 224 bytes."""
    _fields_ = [
        ("projViewMatrix", C.c_float * 16), ("viewMatrix", C.c_float * 16), ("viewMatrixInverseTranspose", C.c_float * 16),
        ("viewport", C.c_int32 * 3), ("linkedListAllocatedPerElement", C.c_uint32), ("alphaMin", C.c_float),
        ("alphaWidth", C.c_float), ("pad", C.c_float * 2),
    ]


class OitStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("fragments", "fragmentsStored", "fragmentsTail", "opaqueFragments",
                                          "trianglesDrawn", "trianglesRejected", "llCounter", "tilePairs",
                                          "kernelLaunches")] + \
               [(n, C.c_float) for n in ("msGeometry", "msClear", "msOpaque", "msColor", "msComposite", "msResolve",
                                         "msFrame", "msExchangeWait")]
assert C.sizeof(OitStats) == 104


assert C.sizeof(SceneData) == 224

_lib = None


def load_library():
    """Loads liboit_b200.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("OIT_B200_LIB", LIB_PATH)  # developer override: alternative builds of the same library
    if not os.path.exists(path):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C vk_order_independent_transparency_b200/csrc). There is no CPU fallback.")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.oit_abi_version.restype = C.c_int
    L.oit_default_config.argtypes = [C.POINTER(OitConfig)]
    L.oit_default_config.restype = None
    L.oit_create.argtypes = [C.POINTER(OitConfig), C.POINTER(vp)]
    L.oit_destroy.argtypes = [vp]
    L.oit_last_error.argtypes = [vp]
    L.oit_last_error.restype = C.c_char_p
    L.oit_get_config.argtypes = [vp, C.POINTER(OitConfig)]
    L.oit_get_dims.argtypes = [vp] + [C.POINTER(C.c_uint32)] * 5
    L.oit_set_scene.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32]
    L.oit_set_scene_device.argtypes = [vp, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32]
    L.oit_scene_sizes.argtypes = [C.POINTER(OitConfig)] + [C.POINTER(C.c_uint32)] * 3
    L.oit_generate_scene.argtypes = [C.POINTER(OitConfig), vp, vp]
    L.oit_default_camera.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float),
                                     C.POINTER(C.c_float), C.c_float, C.c_float, C.POINTER(SceneData)]
    L.oit_render.argtypes = [vp, C.POINTER(SceneData)]
    L.oit_set_scene_data.argtypes = [vp, C.POINTER(SceneData)]
    for f in ("oit_begin_frame", "oit_draw_opaque", "oit_draw_transparent", "oit_composite", "oit_resolve", "oit_synchronize"):
        getattr(L, f).argtypes = [vp]
    L.oit_buffer_size.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t)]
    L.oit_download.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.oit_upload.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.oit_device_ptr.argtypes = [vp, C.c_int]
    L.oit_device_ptr.restype = vp
    L.oit_read_color.argtypes = [vp, vp, C.c_size_t]
    L.oit_get_stats.argtypes = [vp, C.POINTER(OitStats)]
    L.oit_stream.argtypes = [vp]
    L.oit_stream.restype = vp
    L.oit_local_row_to_global.argtypes = [vp, C.c_uint32, C.POINTER(C.c_uint32)]
    L.oit_band_gather_unique_id.argtypes = [vp]
    L.oit_enable_band_gather.argtypes = [vp, vp]
    L.oit_band_peer_export.argtypes = [vp, vp]
    L.oit_band_peer_enable.argtypes = [vp, vp, C.c_uint32]
    L.oit_band_peer_disable.argtypes = [vp]
    L.oit_generate_spheres.argtypes = [C.POINTER(OitConfig), vp]
    L.oit_set_scene_spheres.argtypes = [vp, vp, C.c_uint32, C.c_int32]
    _lib = L
    return L


class State:
    """The reference's State (oit.h:64-116): same field names, encodings and defaults."""

    def __init__(self, algorithm=OIT_SPINLOCK, oitLayers=8, linkedListAllocatedPerElement=10, percentTransparent=100,
                 tailBlend=True, interlockIsOrdered=True, numObjects=1024, subdiv=16, scaleMin=0.1, scaleWidth=0.9,
                 aaType=AA_NONE, sceneStdlib=STDLIB_LIBSTDCXX):
        self.sceneStdlib = sceneStdlib  # harness only: whose std::default_random_engine draws the scene (see oit_b200.h)
        self.algorithm = algorithm
        self.oitLayers = oitLayers
        self.linkedListAllocatedPerElement = linkedListAllocatedPerElement
        self.percentTransparent = percentTransparent
        self.tailBlend = tailBlend
        self.interlockIsOrdered = interlockIsOrdered
        self.numObjects = numObjects
        self.subdiv = subdiv
        self.scaleMin = scaleMin
        self.scaleWidth = scaleWidth
        self.aaType = aaType
        self.recomputeAntialiasingSettings()

    def recomputeAntialiasingSettings(self):
        """oit.h:84-115"""
        table = {AA_NONE: (1, False, 1), AA_MSAA_4X: (4, False, 1), AA_SSAA_4X: (4, True, 1), AA_SUPER_4X: (1, False, 2),
                 AA_MSAA_8X: (8, False, 1), AA_SSAA_8X: (8, True, 1)}
        if self.aaType not in table:
            raise ValueError("Antialiasing mode not implemented!")
        self.msaa, self.sampleShading, self.supersample = table[self.aaType]

    def coverageShading(self):
        return self.msaa > 1 and not self.sampleShading

    def to_config(self, width, height, device=0, bandCount=1, bandIndex=0, stripRows=32):
        cfg = OitConfig()
        cfg.algorithm, cfg.oitLayers = self.algorithm, self.oitLayers
        cfg.linkedListAllocatedPerElement, cfg.percentTransparent = self.linkedListAllocatedPerElement, self.percentTransparent
        cfg.tailBlend, cfg.interlockIsOrdered = int(self.tailBlend), int(self.interlockIsOrdered)
        cfg.numObjects, cfg.subdiv, cfg.scaleMin, cfg.scaleWidth = self.numObjects, self.subdiv, self.scaleMin, self.scaleWidth
        cfg.aaType, cfg.width, cfg.height, cfg.device = self.aaType, width, height, device
        cfg.bandCount, cfg.bandIndex, cfg.stripRows = bandCount, bandIndex, stripRows
        cfg.reserved[1] = int(self.sceneStdlib)
        return cfg


def generate_scene(state):
    """initScene (main.cpp:334-391): returns (vertices[n,10] float32, indices uint32, indicesPerObject)."""
    L = load_library()
    cfg = state.to_config(16, 16)
    nv, ni, ipo = C.c_uint32(), C.c_uint32(), C.c_uint32()
    if L.oit_scene_sizes(C.byref(cfg), C.byref(nv), C.byref(ni), C.byref(ipo)) != 0:
        raise OitError(-1, "bad scene parameters")
    verts = np.empty((nv.value, 10), np.float32)
    idx = np.empty(ni.value, np.uint32)
    L.oit_generate_scene(C.byref(cfg), verts.ctypes.data, idx.ctypes.data)
    return verts, idx, ipo.value


def generate_spheres(state):
    """The per-object table initScene draws (main.cpp:350-369): float32 [numObjects, 8] = centre xyz, radius, colour rgba
    (32 bytes per object instead of the flattened mesh; Sample.setSceneSpheres flattens it on the device)."""
    L = load_library()
    cfg = state.to_config(16, 16)
    table = np.empty((max(int(state.numObjects), 0), 8), np.float32)
    if L.oit_generate_spheres(C.byref(cfg), table.ctypes.data) != 0:
        raise OitError(-1, "bad scene parameters")
    return table


def default_camera(width, height, fov=45.0, eye=(0.0, 0.0, 12.0), center=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), near=0.1, far=100.0):
    """Camera of Sample::onAttach (main.cpp:79-82,121-123): eye (0,0,0.75*GRID_SIZE) looking at the origin, fov 45."""
    L = load_library()
    sd = SceneData()
    f3 = C.c_float * 3
    r = L.oit_default_camera(width, height, fov, f3(*eye), f3(*center), f3(*up), near, far, C.byref(sd))
    if r != 0:
        raise OitError(r, "bad camera parameters")
    return sd


class _DeviceArray:
    """Minimal __cuda_array_interface__ holder so torch / cupy can wrap a library-owned device buffer without a copy."""

    def __init__(self, ptr, shape, typestr, owner):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 2}
        self._owner = owner


class Sample:
    """Host-side mirror of the reference's `Sample` hot-path interface (oit.h:379-425, oitRender.cpp).

    `onRender(ubo)` is the whole frame; the `clearTransparent*` / `drawTransparent*` methods keep the reference's names and
    split so that tests read like the reference's frame recorder."""

    def __init__(self, state, width, height, device=0, bandCount=1, bandIndex=0, stripRows=32, keepIntermediates=False):
        """keepIntermediates: onRender keeps m_colorImage (the per-sample colour target) in device memory so that it can
        be downloaded; by default the frame is fused (colour pass + composite + resolve per tile) and only the A-buffer,
        the aux images and the resolved frame exist afterwards.  The stage-by-stage calls always keep it."""
        self.L = load_library()
        self.state = state
        self.width, self.height = width, height
        self.cfg = state.to_config(width, height, device, bandCount, bandIndex, stripRows)
        self.cfg.reserved[0] = 1 if keepIntermediates else 0
        h = C.c_void_p()
        r = self.L.oit_create(C.byref(self.cfg), C.byref(h))
        if r != 0:
            raise OitError(r, self.L.oit_last_error(None).decode())
        self.h = h
        d = [C.c_uint32() for _ in range(5)]
        self.L.oit_get_dims(self.h, *[C.byref(x) for x in d])
        self.bufW, self.bufH, self.msaa, self.sampleShading, self.localRows = [x.value for x in d]
        self.sampleShading = bool(self.sampleShading)
        self._scene_keepalive = None

    # ---- lifetime -----------------------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            dist = getattr(self, "_peer_dist", None)
            if dist is not None:
                # peer-memory split frame: nobody may still store into a buffer that is about to be unmapped / freed
                self._peer_dist = None
                self.L.oit_synchronize(self.h)
                dist.barrier()
                self.L.oit_band_peer_disable(self.h)
                dist.barrier()
            self.L.oit_destroy(self.h)
            self.h = None

    def __del__(self):
        # Finalizers run at different times on different ranks (or after the process group is gone), so they must not run
        # the collective tear-down of the peer exchange: a peer-enabled Sample has to be closed explicitly (close() or a
        # `with` block).  If it was not, this band unmaps the other bands' buffers and LEAKS its own exported frame buffer --
        # other bands may still have it mapped -- instead of freeing it under them.
        try:
            if getattr(self, "h", None) and getattr(self, "_peer_dist", None) is not None:
                import warnings
                warnings.warn("Sample with the peer-memory exchange enabled was not closed explicitly: its exported frame buffer is leaked",
                              ResourceWarning)
                self._peer_dist = None
                self.L.oit_synchronize(self.h)
                self.L.oit_band_peer_disable(self.h)   # first phase only: local unmap, nothing is freed
                self.h = None
                return
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def _check(self, r):
        if r != 0:
            raise OitError(r, self.L.oit_last_error(self.h).decode())

    # ---- scene ----------------------------------------------------------------------------------------------------
    def initScene(self):
        verts, idx, ipo = generate_scene(self.state)
        self.setScene(verts, idx, ipo)
        return verts, idx, ipo

    def setScene(self, verts, idx, indicesPerObject):
        verts = np.ascontiguousarray(verts, np.float32)
        idx = np.ascontiguousarray(idx, np.uint32)
        self._check(self.L.oit_set_scene(self.h, verts.ctypes.data, verts.shape[0], idx.ctypes.data, idx.size, indicesPerObject))

    def setSceneSpheres(self, table, subdiv=None):
        """Instanced scene input: `table` float32 [n, 8] (generate_spheres); the mesh is flattened on the device."""
        table = np.ascontiguousarray(table, np.float32).reshape(-1, 8)
        self._check(self.L.oit_set_scene_spheres(self.h, table.ctypes.data, table.shape[0],
                                                 int(self.state.subdiv if subdiv is None else subdiv)))

    def setSceneDevice(self, dverts_ptr, nVerts, didx_ptr, nIndices, indicesPerObject, keepalive=None):
        self._scene_keepalive = keepalive
        self._check(self.L.oit_set_scene_device(self.h, dverts_ptr, nVerts, didx_ptr, nIndices, indicesPerObject))

    # ---- frame ------------------------------------------------------------------------------------------------------
    def onRender(self, ubo):
        """Sample::onRender (oitRender.cpp:28-154)."""
        self._check(self.L.oit_render(self.h, C.byref(ubo)))

    def updateUniformBuffer(self, ubo):
        self._check(self.L.oit_set_scene_data(self.h, C.byref(ubo)))

    def _clear(self, *algos):
        if self.state.algorithm not in algos:
            raise ValueError("Algorithm case not called in switch statement!")  # oitRender.cpp:65
        self._check(self.L.oit_begin_frame(self.h))

    def clearTransparentSimple(self):
        self._clear(OIT_SIMPLE)

    def clearTransparentLinkedList(self):
        self._clear(OIT_LINKEDLIST)

    def clearTransparentLoop(self):
        self._clear(OIT_LOOP)

    def clearTransparentLoop64(self):
        self._clear(OIT_LOOP64)

    def clearTransparentLock(self, useInterlock):
        self._clear(OIT_INTERLOCK if useInterlock else OIT_SPINLOCK)

    def beginFrame(self):
        """Clears for whatever the algorithm is (WBOIT's are done by its render pass, oitRender.cpp:60-62)."""
        self._check(self.L.oit_begin_frame(self.h))

    def drawOpaque(self):
        self._check(self.L.oit_draw_opaque(self.h))

    def _draw(self, *algos):
        if self.state.algorithm not in algos:
            raise ValueError("Algorithm case not called in switch statement!")  # oitRender.cpp:147
        self._check(self.L.oit_draw_transparent(self.h))
        self._check(self.L.oit_composite(self.h))

    def drawTransparentSimple(self):
        self._draw(OIT_SIMPLE)

    def drawTransparentLinkedList(self):
        self._draw(OIT_LINKEDLIST)

    def drawTransparentLoop(self):
        self._draw(OIT_LOOP)

    def drawTransparentLoop64(self):
        self._draw(OIT_LOOP64)

    def drawTransparentLock(self, useInterlock):
        self._draw(OIT_INTERLOCK if useInterlock else OIT_SPINLOCK)

    def drawTransparentWeighted(self):
        self._draw(OIT_WEIGHTED)

    def drawTransparentColorOnly(self):
        """Only the colour pass(es), stopping at the fragment barrier (for A-buffer dumps)."""
        self._check(self.L.oit_draw_transparent(self.h))

    def composite(self):
        self._check(self.L.oit_composite(self.h))

    def copyOffscreenToBackBuffer(self):
        self._check(self.L.oit_resolve(self.h))

    def synchronize(self):
        self._check(self.L.oit_synchronize(self.h))

    def saveImage(self, path):
        """saveImageToFile of the test sequencer (main.cpp:887-891): this band's frame as a PNG."""
        save_png(self.readColor(), path)

    # ---- results ------------------------------------------------------------------------------------------------------
    def buffer_size(self, which):
        n = C.c_size_t()
        self._check(self.L.oit_buffer_size(self.h, which, C.byref(n)))
        return n.value

    def download(self, which, dtype=np.uint32):
        n = self.buffer_size(which)
        out = np.empty(n // np.dtype(dtype).itemsize, dtype)
        if n:
            self._check(self.L.oit_download(self.h, which, out.ctypes.data, n))
        return out

    def upload(self, which, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self.L.oit_upload(self.h, which, arr.ctypes.data, arr.nbytes))

    def device_array(self, which, dtype="<u4"):
        ptr = self.L.oit_device_ptr(self.h, which)
        if not ptr:
            return None
        return _DeviceArray(ptr, (self.buffer_size(which) // np.dtype(dtype).itemsize,), dtype, self)

    def readColor(self, out=None):
        """The resolved BGRA8 rows this band owns (m_viewportImage), uint32 [localRows, width]."""
        if out is None:
            out = np.empty((self.localRows, self.width), np.uint32)
        if out.size:
            self._check(self.L.oit_read_color(self.h, out.ctypes.data, out.nbytes))
        return out

    def colorSamples(self):
        return self.download(BUF_COLOR).reshape(-1, self.bufW, self.msaa)

    def stats(self):
        s = OitStats()
        self._check(self.L.oit_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in OitStats._fields_}

    def enableBandGather(self, dist=None):
        """Split frame: make onRender end with the in-library band gather (ONE ncclAllGather + row interleave inside the
        frame graph).  Collective over the bandCount ranks; `dist` = an initialised torch.distributed (any backend), used
        only to hand rank 0's 128-byte NCCL id to the other ranks.  Afterwards frameDevice() is the whole frame."""
        import torch
        ident = (C.c_ubyte * 128)()
        if self.cfg.bandIndex == 0:
            self._check(self.L.oit_band_gather_unique_id(ident))
        if self.cfg.bandCount > 1:
            if dist is None:
                import torch.distributed as dist
            box = [bytes(ident)]
            dist.broadcast_object_list(box, src=0)
            ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        self._check(self.L.oit_enable_band_gather(self.h, ident))

    def enableBandPeers(self, dist=None):
        """Split frame over NVLink peer memory: the frame kernel of onRender stores its resolved pixels straight into the
        frame buffers of EVERY band (CUDA IPC mappings; two per band, frames alternate), and two flag rounds per frame keep the
        bands in step one frame late (synchronize() / readFrame() / frameDevice() after synchronize() see the latest frame).  Collective over
        the bandCount ranks; `dist` = an initialised torch.distributed, used only to exchange the 64-byte IPC handles.
        Returns False (on every rank, nothing enabled) when some rank cannot map its peers: use enableBandGather then."""
        if dist is None:
            import torch.distributed as dist
        world = max(self.cfg.bandCount, 1)
        handle = (C.c_ubyte * 64)()
        ok = self.L.oit_band_peer_export(self.h, handle) == 0
        box = [None] * world
        dist.all_gather_object(box, (ok, bytes(handle)))
        ok = all(b[0] for b in box)
        if ok:
            blob = b"".join(b[1] for b in box)
            ok = self.L.oit_band_peer_enable(self.h, (C.c_ubyte * len(blob)).from_buffer_copy(blob), world) == 0
        box = [None] * world
        dist.all_gather_object(box, ok)
        if all(box):
            self._peer_dist = dist
            return True
        self.L.oit_band_peer_disable(self.h)  # unmap whatever was mapped ...
        dist.barrier()
        self.L.oit_band_peer_disable(self.h)  # ... then release the exported buffer
        return False

    def frameDevice(self):
        """The gathered full frame (uint32 BGRA8 [height, width]) as a __cuda_array_interface__ object (zero copy).  Completes
        the latest frame first (with the peer exchange the frames alternate between two buffers, and the pointer is the one
        of the latest COMPLETED frame); valid until the next onRender."""
        self.synchronize()
        ptr = self.L.oit_device_ptr(self.h, BUF_FRAME)
        return _DeviceArray(ptr, (self.height, self.width), "<i4", self) if ptr else None

    def readFrame(self):
        return self.download(BUF_FRAME).reshape(self.height, self.width)

    def localRowToGlobal(self, r):
        g = C.c_uint32()
        self._check(self.L.oit_local_row_to_global(self.h, r, C.byref(g)))
        return g.value

    def globalRows(self):
        """Global output row of every local row of this band."""
        strip = self.cfg.stripRows or 32
        r = np.arange(self.localRows)
        return ((r // strip) * max(self.cfg.bandCount, 1) + self.cfg.bandIndex) * strip + r % strip


def bgra_to_rgba_image(final):
    b = np.ascontiguousarray(final).view(np.uint8).reshape(final.shape[0], final.shape[1], 4)
    return b[..., [2, 1, 0, 3]].copy()


def save_png(final, path):
    """What the reference's test sequencer does with m_viewportImage after every sequence (saveImageToFile,
    main.cpp:887-891): the BGRA8 frame (`Sample.readColor()` / `readFrame()`) as an 8-bit RGBA PNG.  Dependency-free
    (zlib + struct); the bytes are the sRGB-encoded values the frame holds, written unchanged."""
    import struct
    import zlib
    rgba = bgra_to_rgba_image(final)
    h, w = rgba.shape[:2]
    raw = np.concatenate([np.zeros((h, 1), np.uint8), rgba.reshape(h, w * 4)], axis=1).tobytes()   # filter type 0 per row

    def chunk(tag, data):
        body = tag + data
        return struct.pack(">I", len(data)) + body + struct.pack(">I", zlib.crc32(body) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 6, 0, 0, 0))
                + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))
