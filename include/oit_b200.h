/*
 * oit_b200.h -- C ABI of the B200-native order-independent-transparency library (liboit_b200.so).
 *
 * Drop-in boundary for the hot path of nvpro-samples/vk_order_independent_transparency: the stages that
 * Sample::onRender records between "clear" and "copyOffscreenToBackBuffer" (oitRender.cpp:28-154), i.e.
 *
 *     clearTransparent{Simple,LinkedList,Loop,Loop64,Lock}      oit.h:379-425, oitRender.cpp:156-356
 *     the opaque draw                                           oitRender.cpp:113-122
 *     drawTransparent{Simple,LinkedList,Loop,Loop64,Lock,Weighted}  (colour pass + barrier + composite)
 *     copyOffscreenToBackBuffer (MSAA resolve / 2x downsample)  main.cpp:645-774
 *
 * Plain C: pointers, sizes and PODs only; every call returns 0 on success or a negative OitResult and never
 * throws or aborts across the boundary (the reference aborts through NVVK_CHECK / assert, oitRender.cpp:65,147).
 * One host thread drives one context; a context owns one CUDA device, one stream and all device memory.
 * There is NO CPU fallback: without a CUDA device every entry point fails with OIT_ERR_CUDA.
 */
#ifndef OIT_B200_H
#define OIT_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OIT_B200_ABI_VERSION 1

/* algorithm / antialiasing encodings: identical to shaders/common.h:44-63 */
enum
{
  OIT_SIMPLE     = 0,
  OIT_LINKEDLIST = 1,
  OIT_LOOP       = 2,
  OIT_LOOP64     = 3,
  OIT_SPINLOCK   = 4,
  OIT_INTERLOCK  = 5,
  OIT_WEIGHTED   = 6,
  OIT_NUM_ALGORITHMS
};
enum
{
  OIT_AA_NONE     = 0,
  OIT_AA_MSAA_4X  = 1, /* 4 samples, per-pixel shading, coverage masks in the A-buffer */
  OIT_AA_SSAA_4X  = 2, /* 4 samples, per-sample shading, A-buffer x4 */
  OIT_AA_SUPER_4X = 3, /* render at 2W x 2H, LINEAR downsample */
  OIT_AA_MSAA_8X  = 4,
  OIT_AA_SSAA_8X  = 5,
  OIT_NUM_AATYPES
};

typedef enum OitResult
{
  OIT_OK              = 0,
  OIT_ERR_INVALID_ARG = -1, /* unknown enum, size 0, layers outside {1..32}, null pointer ... */
  OIT_ERR_CUDA        = -2, /* no device / a CUDA call failed; oit_last_error() has the CUDA string */
  OIT_ERR_NO_SCENE    = -3, /* draw before oit_set_scene */
  OIT_ERR_OUT_OF_MEMORY = -4,
  OIT_ERR_SIZE        = -5, /* host buffer size does not match the device buffer */
  OIT_ERR_UNSUPPORTED = -6
} OitResult;

/*
 * Parameter surface = the reference's State (oit.h:64-82): same names, same integer encodings, same defaults
 * (see oit_default_config).  The fields after aaType are what the reference takes from its window
 * (width/height) and what the split-frame multi-GPU mode adds.
 */
typedef struct OitConfig
{
  uint32_t algorithm;                     /* default OIT_SPINLOCK */
  uint32_t oitLayers;                     /* OIT_LAYERS, 1..32 (GUI offers 1,2,4,8,16,32: oitGui.cpp:257-270) */
  int32_t  linkedListAllocatedPerElement; /* linked-list pool = N*W*H[*msaa] nodes (oit.cpp:119-126,148-152) */
  int32_t  percentTransparent;            /* 0..100; first spheres transparent, last opaque (oitRender.cpp:68-78) */
  uint32_t tailBlend;
  uint32_t interlockIsOrdered;            /* both values use the primitive-ordered critical section here */
  int32_t  numObjects;                    /* only used by oit_generate_scene */
  int32_t  subdiv;
  float    scaleMin;
  float    scaleWidth;
  uint32_t aaType;
  uint32_t width;                         /* viewport size before the supersample factor */
  uint32_t height;
  int32_t  device;                        /* CUDA device ordinal */
  /* sort-first split frame: the context renders only the row strips it owns.  Strip k (stripRows rows of the
     output image) belongs to band k % bandCount.  bandCount = 1 renders the whole frame. */
  uint32_t bandCount;
  uint32_t bandIndex;
  uint32_t stripRows;                     /* multiple of 16; 0 = default (32) */
  uint32_t reserved[4];                   /* [0] bit 0: keep the intermediate images (staged frame); [1]: OIT_CFG_SCENE_STDLIB */
} OitConfig;

/* Scene generator only: std::default_random_engine / uniform_real_distribution<float> (main.cpp:350-351) are
   implementation-defined.  libstdc++: minstd_rand0, one draw / (2^31 - 2); MSVC's STL: mt19937, one draw / 2^32 -- the
   build that produced the screenshot in the reference's doc/ directory (tests/test_reference_screenshot.py). */
#define OIT_STDLIB_LIBSTDCXX 0u
#define OIT_STDLIB_MSVC 1u
#define OIT_CFG_SCENE_STDLIB(cfg) ((cfg)->reserved[1])

/* shaderio::SceneData, std140, 224 bytes (shaders/common.h:77-92); matrices column-major like glm.
   viewport and linkedListAllocatedPerElement are overwritten by the library exactly as
   updateUniformBuffer (main.cpp:628-637) and createFrameImages (oit.cpp:102-152) do. */
typedef struct OitSceneData
{
  float    projViewMatrix[16];
  float    viewMatrix[16];
  float    viewMatrixInverseTranspose[16];
  int32_t  viewport[3];
  uint32_t linkedListAllocatedPerElement;
  float    alphaMin;
  float    alphaWidth;
  float    pad[2];
} OitSceneData;

/* stage names follow the reference's profiler sections (oitRender.cpp:158-389) */
typedef struct OitStats
{
  uint64_t fragments;          /* colour-pass invocations of the transparent draw: the metric's F */
  uint64_t fragmentsStored;
  uint64_t fragmentsTail;
  uint64_t opaqueFragments;
  uint64_t trianglesDrawn;
  uint64_t trianglesRejected;  /* entirely behind the near plane, or not representable (far plane, guard band);
                                  triangles that CROSS the near plane are clipped, not rejected */
  uint64_t llCounter;          /* linked list: final counter value (may exceed the pool) */
  uint64_t tilePairs;          /* (tile, triangle) pairs binned this frame */
  uint64_t kernelLaunches;     /* kernels launched by the last oit_render */
  float    msGeometry;         /* vertex transform + triangle setup + binning */
  float    msClear;            /* <Tech>Clear + colour/depth clear */
  float    msOpaque;
  float    msColor;            /* <Tech>Color (+ LoopDepth) */
  float    msComposite;        /* <Tech>Composite */
  float    msResolve;          /* copyOffscreenToBackBuffer */
  float    msFrame;            /* whole oit_render on the device */
  float    msExchangeWait;     /* split frame over peer memory: time this band waited for the other bands (READY + DONE rounds) */
} OitStats;

/* device/host buffers addressable through oit_download / oit_upload / oit_device_ptr */
typedef enum OitBuffer
{
  OIT_BUF_ABUFFER  = 0, /* layout per technique exactly as oit.cpp:84-163 / the shaders index it */
  OIT_BUF_AUX      = 1, /* imgAux      R32UI W x H x layers */
  OIT_BUF_AUXSPIN  = 2, /* imgSpin */
  OIT_BUF_AUXDEPTH = 3, /* imgDepth */
  OIT_BUF_COUNTER  = 4, /* imgCounter  1 x 1 */
  OIT_BUF_COLOR    = 5, /* m_colorImage: BGRA8 sRGB words, [y][x][sample] */
  OIT_BUF_DEPTH    = 6, /* m_depthImage: float [y][x][sample]; only allocated when something opaque is drawn */
  OIT_BUF_WACCUM   = 7, /* WBOIT RGBA16F [y][x][sample][4] */
  OIT_BUF_WREVEAL  = 8, /* WBOIT R16F   [y][x][sample] */
  OIT_BUF_FINAL    = 9, /* m_viewportImage: BGRA8, width x (rows owned by this band), sRGB-encoded bytes */
  OIT_BUF_FRAME    = 10 /* split frame with the band gather enabled: the whole m_viewportImage, width x height, on every rank */
} OitBuffer;

typedef struct OitCtx OitCtx;

/* ---- lifetime ---------------------------------------------------------------------------------------- */
int         oit_abi_version(void);
/* host-only self check (no GPU needed): the sRGB8 encoder of the frame kernels (bucket table + one threshold) against its
   definition (largest code whose threshold the value has reached) on every `stride`-th float of [0, 1] and on the special
   values; stride 1 checks all 2^30 + 2^23 + 1 floats.  Returns 0; *mismatches must be 0. */
int         oit_selfcheck_srgb_encoder(uint32_t stride, uint64_t* checked, uint64_t* mismatches);
void        oit_default_config(OitConfig* cfg);                 /* State{} defaults, 1280x720, 1 band */
int         oit_create(const OitConfig* cfg, OitCtx** out);     /* = updateRendererFromState(true,true), main.cpp:130-250 */
int         oit_destroy(OitCtx* ctx);
const char* oit_last_error(const OitCtx* ctx);                  /* ctx may be NULL: error of the last failed create */
int         oit_get_config(const OitCtx* ctx, OitConfig* out);
/* derived sizes: render-target width/height (after supersample), msaa, sampleShading, rows owned by this band */
int         oit_get_dims(const OitCtx* ctx, uint32_t* bufW, uint32_t* bufH, uint32_t* msaa, uint32_t* sampleShading,
                         uint32_t* localRows);

/* ---- scene: what initScene uploads (main.cpp:334-417) ------------------------------------------------------ */
/* host pointers, copied.  Vertex = pos3f, normal3f, colour4f, 40-byte stride (utilities_vk.h:48-61) */
int oit_set_scene(OitCtx* ctx, const void* vertices, uint32_t nVerts, const uint32_t* indices, uint32_t nIndices,
                  uint32_t indicesPerObject);
/* device pointers on ctx's device, NOT copied; the caller keeps them alive */
int oit_set_scene_device(OitCtx* ctx, const void* dVertices, uint32_t nVerts, const uint32_t* dIndices, uint32_t nIndices,
                         uint32_t indicesPerObject);
/* host-side scene generator + camera of the sample (harness) */
int oit_scene_sizes(const OitConfig* cfg, uint32_t* nVerts, uint32_t* nIndices, uint32_t* indicesPerObject);
int oit_generate_scene(const OitConfig* cfg, void* vertices, uint32_t* indices);
int oit_default_camera(uint32_t width, uint32_t height, float fovDeg, const float eye[3], const float center[3],
                       const float up[3], float zNear, float zFar, OitSceneData* out);

/* ---- instanced scene input: the sphere cloud as a 32-byte-per-object table instead of the flattened mesh.
   initScene (main.cpp:346-391) draws centre, radius and colour per object and flattens numObjects copies of one UV sphere
   on the host; here the host hands over only the table and the flattening runs on the device (k_expand_spheres), with the
   same arithmetic (pos = unit * radius + centre, separate multiply and add), so the vertex / index buffers -- and every
   result -- are bit-identical to oit_generate_scene + oit_set_scene.  34.8 MB of upload become 32 KB for the default
   scene. */
typedef struct OitSphere
{
  float center[3];
  float radius;
  float color[4]; /* rgb already squared (main.cpp:366-369), alpha */
} OitSphere;
/* the table oit_generate_scene flattens: cfg->numObjects entries (numObjects, scaleMin, scaleWidth are read) */
int oit_generate_spheres(const OitConfig* cfg, OitSphere* spheres);
/* host pointer, copied; subdiv as in State (2..1024); nSpheres * vertices per sphere must fit 32 bits */
int oit_set_scene_spheres(OitCtx* ctx, const OitSphere* spheres, uint32_t nSpheres, int32_t subdiv);

/* ---- frame -------------------------------------------------------------------------------------------- */
/* Sample::onRender (oitRender.cpp:28-154): clear, opaque, transparent colour pass(es), composite, resolve.
   Asynchronous, like recording + submitting the frame's command buffer: the frame is enqueued on the context's stream
   (one CUDA graph launch) and may still be running when the call returns; up to 4 frames can be in flight.  A frame is
   COMPLETED -- waited for, and rendered again if one of the library's internal buffers had to grow -- by oit_synchronize,
   oit_download / oit_read_color, oit_get_stats, by every scene change and by the stage-by-stage calls below.  Hosts that
   consume oit_device_ptr() memory directly call oit_synchronize first.  OIT_B200_SYNC_RENDER=1 in the environment makes
   oit_render complete its frame before returning. */
int oit_render(OitCtx* ctx, const OitSceneData* ubo);
/* the same stage by stage, for tests and for hosts that interleave their own work (each is asynchronous on the
   context's stream; oit_synchronize waits) */
int oit_set_scene_data(OitCtx* ctx, const OitSceneData* ubo);   /* updateUniformBuffer */
int oit_begin_frame(OitCtx* ctx);          /* vertex stage + binning + clearTransparent* + render-pass clears */
int oit_draw_opaque(OitCtx* ctx);          /* oitRender.cpp:113-122 */
int oit_draw_transparent(OitCtx* ctx);     /* colour pass(es) of drawTransparent*, up to the fragment barrier */
int oit_composite(OitCtx* ctx);            /* the full-screen composite draw of drawTransparent* */
int oit_resolve(OitCtx* ctx);              /* copyOffscreenToBackBuffer */
int oit_synchronize(OitCtx* ctx);

/* ---- results / dumps ------------------------------------------------------------------------------------- */
int   oit_buffer_size(const OitCtx* ctx, OitBuffer which, size_t* bytes);
int   oit_download(OitCtx* ctx, OitBuffer which, void* host, size_t bytes);   /* A-buffer dump etc. */
int   oit_upload(OitCtx* ctx, OitBuffer which, const void* host, size_t bytes); /* composite-from-dump */
void* oit_device_ptr(OitCtx* ctx, OitBuffer which);                          /* NULL if not allocated */
/* convenience == oit_download(OIT_BUF_FINAL): BGRA8 rows owned by this band, top to bottom */
int   oit_read_color(OitCtx* ctx, void* bgra8, size_t bytes);
int   oit_get_stats(OitCtx* ctx, OitStats* out);
void* oit_stream(OitCtx* ctx);             /* the cudaStream_t the context launches on */
/* global row index of local row r of this band (for reassembling the gathered strips) */
int   oit_local_row_to_global(const OitCtx* ctx, uint32_t localRow, uint32_t* globalRow);

/* ---- split frame: band gather inside the library (no reference counterpart: the sample is single GPU) -------------------
   One process per GPU, each with a context created with bandCount = world size and bandIndex = rank.  Rank 0 calls
   oit_band_gather_unique_id and distributes the 128 bytes (any host transport; the Python mirror uses torch.distributed);
   then EVERY rank calls oit_enable_band_gather (collective).  From then on oit_render ends with ONE ncclAllGather over
   NVLink of the resolved strips (in place) + the row interleave, captured into the frame graph, and OIT_BUF_FRAME holds
   the whole frame on every rank.  width must be a multiple of 4.  NCCL is loaded with dlopen("libnccl.so.2"). */
int oit_band_gather_unique_id(void* id128);
int oit_enable_band_gather(OitCtx* ctx, const void* id128);

/* ---- split frame over NVLink peer memory (preferred; same contexts as above, no reference counterpart) ----------------
   Every band owns two whole-frame buffers (frames alternate between them; OIT_BUF_FRAME is the one that holds the latest
   completed frame).  oit_band_peer_export allocates them and returns their CUDA IPC handle (64 bytes); the host gathers the
   handles of all bands in band order (any transport) and passes them to oit_band_peer_enable, which maps the other bands'
   buffers.  From then on the frame kernel of oit_render stores its resolved pixels straight into ALL bands' frame buffers
   while it renders (no collective after the frame; from four bands up a few CTAs of the frame kernel do nothing but push
   finished tiles over NVLink), and two flag rounds per frame (device-side sequence numbers, part of the frame graph) keep
   the bands in step ONE FRAME LATE, so that a stream of oit_render calls never idles at a barrier; the calls that complete a
   frame (oit_synchronize, oit_download, oit_get_stats, ...) wait until every band's strips of the latest frame have arrived.
   OIT_BUF_FRAME / oit_device_ptr(OIT_BUF_FRAME) then is the whole frame on every band, valid until the next oit_render
   (ask for the pointer again after every completed frame: it alternates).  Every band must call oit_render the same number
   of times.
   Tear-down: all bands finish rendering, host barrier, oit_band_peer_disable on every band (unmaps the other bands'
   buffers, frees nothing), host barrier, oit_destroy -- or a second oit_band_peer_disable, which releases the exported
   buffer and leaves the context usable (e.g. to fall back to the NCCL gather when some band could not map its peers).
   Returns OIT_ERR_UNSUPPORTED when the devices cannot map each other's memory (fall back to the NCCL gather above). */
int oit_band_peer_export(OitCtx* ctx, void* handle64);
int oit_band_peer_enable(OitCtx* ctx, const void* handles, uint32_t count);
int oit_band_peer_disable(OitCtx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* OIT_B200_H */
