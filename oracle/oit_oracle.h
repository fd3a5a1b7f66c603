/*
 * oit_oracle.h -- C API of the CPU ORACLE for the order-independent-transparency hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (vk_order_independent_transparency_b200/,
 * include/) may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker / the CPU baseline.
 *
 * PINNING: the reference (nvpro-samples/vk_order_independent_transparency) cannot be built here (needs
 * Vulkan, nvpro_core2, shaderc) and its tests hold no golden vectors (test.py only checks the exit code,
 * main.cpp:887-891 writes PNGs that are never compared).  The oracle is pinned against
 *  (a) the ONE output the reference publishes, doc/vk_order_independent_transparency.png (README.md:5):
 *      Interlock, 16 layers, MSAA 4x, tail blend, default scene from an MSVC build (mt19937), default
 *      camera, 1920x1017 -- reproduced with mean |diff| 0.13/255, 71 % identical pixels
 *      (tests/test_reference_screenshot.py; tolerance stated there);
 *  (b) the README's worked examples (README.md:45,53-66,72,80,86-96);
 *  (c) the libstdc++ random-number known answers for the scene generator (main.cpp:350-369).
 * Below image level (A-buffer words, counters, the other techniques' intermediates) PARITY IS UNPINNED:
 * it rests on this restatement of the GLSL.  See DESIGN.md "Oracle".
 */
#ifndef OIT_ORACLE_H
#define OIT_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same field names / integer encodings as the reference's State (oit.h:64-82) + target size. */
typedef struct OracleConfig
{
  uint32_t algorithm;                     /* OIT_* (shaders/common.h:44-51) */
  uint32_t oitLayers;                     /* OIT_LAYERS */
  int32_t  linkedListAllocatedPerElement; /* N: pool = N*W*H[*msaa] nodes (oit.cpp:119-126,148-152) */
  int32_t  percentTransparent;
  uint32_t tailBlend;
  uint32_t interlockIsOrdered;
  int32_t  numObjects;
  int32_t  subdiv;
  float    scaleMin;
  float    scaleWidth;
  uint32_t aaType;                        /* AA_* (shaders/common.h:57-63) */
  uint32_t width;                         /* viewport size BEFORE the supersample factor */
  uint32_t height;
} OracleConfig;

/* shaderio::SceneData, std140, 224 bytes (shaders/common.h:77-92). Matrices column-major (glm). */
typedef struct OracleSceneData
{
  float    projViewMatrix[16];
  float    viewMatrix[16];
  float    viewMatrixInverseTranspose[16];
  int32_t  viewport[3];
  uint32_t linkedListAllocatedPerElement;
  float    alphaMin;
  float    alphaWidth;
  float    pad[2];
} OracleSceneData;

typedef struct OracleStats
{
  uint64_t fragments;        /* colour-pass invocations of the transparent draw (metric F) */
  uint64_t fragmentsStored;  /* invocations that ended up in the A-buffer (did not tail blend) */
  uint64_t fragmentsTail;    /* invocations that emitted a non-zero ROP colour */
  uint64_t opaqueFragments;
  uint64_t trianglesDrawn;   /* transparent triangles submitted */
  uint64_t trianglesRejected;/* entirely behind the near plane, or not representable (far plane, guard band) */
  uint64_t llCounter;        /* final value of the linked-list counter */
} OracleStats;

typedef struct OracleCtx OracleCtx;

/* scene + camera harness (main.cpp:334-391, 79-82,121-123,625-637) */
int  oracle_scene_sizes(const OracleConfig* cfg, uint32_t* nVerts, uint32_t* nIndices, uint32_t* indicesPerObject);
int  oracle_generate_scene(const OracleConfig* cfg, float* vertices /*10 floats each*/, uint32_t* indices);
/* stdlib: whose std::default_random_engine / uniform_real_distribution -- 0 libstdc++ (minstd_rand0), 1 MSVC (mt19937) */
int  oracle_generate_scene_ex(const OracleConfig* cfg, int stdlib, float* vertices, uint32_t* indices);
void oracle_camera(uint32_t width, uint32_t height, float fovDeg, const float eye[3], const float center[3],
                   const float up[3], float zNear, float zFar, OracleSceneData* out);
float oracle_rand_canonical(uint64_t* state); /* one draw of minstd_rand0 + uniform_real_distribution<float> */

OracleCtx* oracle_create(const OracleConfig* cfg);
void       oracle_destroy(OracleCtx*);
int        oracle_set_threads(OracleCtx*, int nThreads); /* 1 = the defining sequential schedule */
int        oracle_set_scene(OracleCtx*, const float* vertices, uint32_t nVerts, const uint32_t* indices,
                            uint32_t nIndices, uint32_t indicesPerObject);
int        oracle_set_scene_data(OracleCtx*, const OracleSceneData* ubo);

/* frame = clear + begin pass + opaque + transparent colour pass(es) + composite + resolve (oitRender.cpp:28-154) */
int oracle_render(OracleCtx*, const OracleSceneData* ubo);
/* the same, stage by stage */
int oracle_begin_frame(OracleCtx*);        /* clearTransparent* + colour/depth clear (oitRender.cpp:43-66,89-91) */
int oracle_draw_opaque(OracleCtx*);        /* oitRender.cpp:113-122 */
int oracle_draw_transparent(OracleCtx*);   /* colour pass(es) of drawTransparent*, up to the barrier */
int oracle_composite(OracleCtx*);          /* the full-screen composite draw */
int oracle_resolve(OracleCtx*);            /* copyOffscreenToBackBuffer (main.cpp:645-774) */

/* one hand-fed invocation of the colour pass (README known-answer tests). rgba = unpremultiplied linear colour.
   pass: 0 = depth pass (Loop32 only), 1 = colour pass.  outColor receives what goes to the ROP. */
int oracle_debug_invoke(OracleCtx*, int pass, uint32_t x, uint32_t y, uint32_t sampleID, uint32_t coverageMask,
                        const float rgba[4], float depth, float viewZ, float outColor[4]);

/* buffer access (pointers stay valid until destroy) */
uint32_t* oracle_abuffer(OracleCtx*, size_t* nWords);
uint32_t* oracle_aux(OracleCtx*, int which /*0 aux,1 spin,2 depth,3 counter*/, size_t* nWords);
uint32_t* oracle_color_samples(OracleCtx*, size_t* nWords); /* BGRA8 [y][x][s] of the (supersampled) target */
float*    oracle_depth_samples(OracleCtx*, size_t* nWords);
uint16_t* oracle_weighted(OracleCtx*, int which /*0 accum half4,1 reveal half*/, size_t* nHalfs);
uint32_t* oracle_final(OracleCtx*, size_t* nWords);          /* BGRA8 width*height */
int       oracle_get_stats(OracleCtx*, OracleStats* out);
int       oracle_buffer_dims(OracleCtx*, uint32_t* bufW, uint32_t* bufH, uint32_t* msaa, uint32_t* sampleShading);

/* small pure functions exposed for table / identity tests */
uint32_t oracle_srgb_encode8(float linear);
float    oracle_srgb_decode8(uint32_t v);
uint32_t oracle_pack_color(const float rgbaLinearUnpremult[4]);
uint16_t oracle_float_to_half(float f);
float    oracle_half_to_float(uint16_t h);
uint32_t oracle_rop_blend(uint32_t dstBGRA8, const float srcPremult[4]);

#ifdef __cplusplus
}
#endif
#endif
