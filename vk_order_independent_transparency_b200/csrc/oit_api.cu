// oit_api.cu -- host side of liboit_b200.so: context, device memory, and the frame skeleton of Sample::onRender
// (oitRender.cpp:28-154) expressed as kernel launches on one CUDA stream.  See include/oit_b200.h for the contract.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "oit_internal.h"
#include "oit_scene.h"

using namespace oit;

namespace {
thread_local std::string g_createError;

double srgbToLinearD(double c) { return c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4); }

// The device's SrgbTables (oit_device.cuh): dec[256] sRGB8 code -> linear, thr[260] encode thresholds (thr[0] = -inf,
// thr[256..] = +inf), a255[256] = v / 255, and the encoder's bucket table.  Returns false if a bucket holds two thresholds.
constexpr int TAB_THR = 256, TAB_A255 = 516;
uint32_t hostEnc8(const float* t, float c)
{
  uint32_t k = 0;
  for(uint32_t step = 128; step; step >>= 1)
    if(c >= t[TAB_THR + k + step])
      k += step;
  return k;
}
bool buildTables(unsigned char* bytes)
{
  float* t = reinterpret_cast<float*>(bytes);
  for(int v = 0; v < 256; v++)
    t[v] = (float)srgbToLinearD(v / 255.0);
  t[TAB_THR] = -std::numeric_limits<float>::infinity();
  for(int k = 1; k < 256; k++)
    t[TAB_THR + k] = (float)srgbToLinearD((k - 0.5) / 255.0);
  for(int k = 256; k < 260; k++)
    t[TAB_THR + k] = std::numeric_limits<float>::infinity();
  for(int v = 0; v < 256; v++)
    t[TAB_A255 + v] = (float)v / 255.0f;
  unsigned char* bucket = bytes + SRGB_TABLE_FLOATS * 4;
  memset(bucket, 0, SRGB_BUCKET_BYTES);
  bool ok = true;
  for(uint32_t i = 0; i < SRGB_BUCKET_COUNT; i++)
  {
    const uint32_t loBits = (SRGB_BUCKET_BASE + i) << 16, hiBits = i + 1 < SRGB_BUCKET_COUNT ? ((SRGB_BUCKET_BASE + i + 1) << 16) - 1u : loBits;
    float          lo, hi;
    memcpy(&lo, &loBits, 4);
    memcpy(&hi, &hiBits, 4);
    const uint32_t kLo = hostEnc8(t, lo), kHi = hostEnc8(t, hi);
    bucket[i]          = (unsigned char)kLo;
    ok                 = ok && kHi <= kLo + 1u;
  }
  return ok && bucket[0] == 0 && bucket[SRGB_BUCKET_COUNT - 1] == 255;
}

struct DevBuf
{
  void*  p     = nullptr;
  size_t bytes = 0;
};

enum EventId
{
  EV_START = 0,
  EV_GEOM,
  EV_CLEAR,
  EV_OPAQUE,
  EV_COLOR,
  EV_COMPOSITE,
  EV_RESOLVE,
  EV_RASTER_START,  // pipelined frames: the raster half starts here (the geometry half of the NEXT frame may overlap it)
  NUM_EVENTS
};
}  // namespace

constexpr int UBO_RING = 4;

struct OitCtx
{
  OitConfig    cfg{};
  int          msaa = 1, supersample = 1;
  bool         sampleShading = false, coverage = false;
  uint32_t     bufW = 0, bufH = 0, localBufH = 0, localOutH = 0;
  uint32_t     stripRows = 32;
  std::string  error;
  cudaStream_t stream = nullptr;
  cudaEvent_t  ev[NUM_EVENTS]{};
  bool         evRecorded[NUM_EVENTS]{};
  FrameParams  fp{};
  OitSceneData ubo{};
  bool         haveUbo = false;
  // buffers
  DevBuf abuf, aux, spin, adepth, counter, color, depth, wacc, wrev, fin, tables, rowLocal;
  // Frame pipeline: everything the GEOMETRY half of a frame produces (post-transform vertices, tile lists, clip table, the
  // per-frame UBO copy, the statistics block) exists twice.  oit_render alternates between the two sets and issues the
  // geometry half on its own high-priority stream, so that the vertex stage + binning of frame n+1 run while the raster
  // half of frame n still occupies the SMs (what a GPU's geometry and pixel pipelines do across draws).
  DevBuf       stats[2], tv[2], uboDev[2];
  int          par       = 0;      // the set the next / current frame uses
  bool         pipelined = true;   // OIT_B200_NO_PIPELINE=1: one set, one stream
  int          experimentFrames = 0;
  cudaStream_t geoStream = nullptr;
  cudaEvent_t  evGeoDone[2]{}, evRasterDone[2]{}, evMain = nullptr;
  bool         rasterRecorded[2]{};
  bool         mainDirty = false;  // work was put on the main stream outside oit_render (scene upload, stage-by-stage calls):
                                   // the next geometry half must not overtake it
  // scene
  DevBuf   verts, indices;
  bool     sceneOwned = false;
  uint32_t nVerts = 0, nIndices = 0, idxPerObj = 0;
  // binning (0 = transparent draw, 1 = opaque draw)
  BinBuffers bins[2][2]{};  // [set][draw]
  uint32_t   pairTotal[2]{};
  uint32_t   drawTris[2]{};
  // per-frame UBO in device memory + its pinned staging copy; the captured frame graph
  DeviceUbo*      hostUbo    = nullptr;  // pinned staging ring of UBO_RING entries (frames may be in flight)
  cudaEvent_t     uboEv[UBO_RING]{};     // recorded after the upload out of slot i
  int             uboSlot      = 0;
  bool            framePending = false;  // oit_render enqueued a frame whose overflow check has not run yet
  bool            asyncRender  = true;   // oit_render returns once the frame is enqueued (OIT_B200_SYNC_RENDER=1: waits)
  // the captured frame, per set: [set][0] = geometry half (pipelined mode only), [set][1] = raster half / the whole frame
  cudaGraph_t     graph[2][2]{};
  cudaGraphExec_t graphExec[2][2]{};
  bool            graphValid = false;
  bool            graphBuilt[2]{};
  uint64_t        graphLaunchCount[2]{};
  bool            useGraph   = true;
  bool            capturing  = false;  // a frame is being issued without host synchronisation
  bool            fuseFrame  = false;  // oit_render: colour pass + composite + resolve in one kernel
  unsigned long long* hostMirror  = nullptr;  // pinned: statistics + pair counts copied back by the frame itself
  bool                mirrorValid = false;
  // split frame: band gather inside the library (oit_gather.cu)
  BandGatherState* gather     = nullptr;
  DevBuf           gatherBuf, frame;
  uint32_t         padRows    = 0;
  bool             gatherWarm = false;  // NCCL has run once outside a capture (connection set-up must not be captured)
  bool             finOwned   = true;   // false once `fin` is this rank's slice of the gather buffer
  // split frame over peer memory (oit_peer.cu): the frame kernel stores into every band's frame buffer
  // instanced scene input (oit_set_scene_spheres): the object table and the unit-sphere template of `sphSubdiv`
  DevBuf sphTable, sphUnitPos, sphUnitTri;
  int    sphSubdiv = 0;
  PeerState* peers     = nullptr;
  bool       peersOpen = false;
  unsigned   peerSeq   = 0;  // frames issued with the exchange: the latest one is in frame buffer peerSeq & 1
  bool       peersUnmapped = false;  // oit_band_peer_disable has run once: the second call frees the exported buffer
  DevBuf     pushQueue;              // pusher CTAs of the linked-list frame kernel (FrameParams::pushQueue)
  int        pushers = -1;           // OIT_B200_PUSHERS; default: 16 from four bands up, otherwise 0 = the tile CTAs store to the
                                     // peers themselves (measured on 8 B200s: colour pass 0.259 -> 0.227 ms with pushers at 8 bands,
                                     // 0.784 -> 0.818 ms at 2 bands, where the remote traffic is small and the pushers' slots are missed)
  int        sortedBuf[2][2]{};  // [set][draw]
  uint32_t*  hostScalar = nullptr;  // pinned
  OitStats   lastStats{};
  uint64_t   launches = 0;
};

namespace {

int fail(OitCtx* c, int code, const std::string& msg)
{
  if(c)
    c->error = msg;
  else
    g_createError = msg;
  return code;
}
#define CUDA_TRY(ctx, call)                                                                                                      \
  do                                                                                                                             \
  {                                                                                                                              \
    cudaError_t e__ = (call);                                                                                                    \
    if(e__ != cudaSuccess)                                                                                                       \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? OIT_ERR_OUT_OF_MEMORY : OIT_ERR_CUDA,                                   \
                  std::string(#call) + ": " + cudaGetErrorString(e__));                                                          \
  } while(0)

int devAlloc(OitCtx* c, DevBuf& b, size_t bytes)
{
  if(b.p)
  {
    cudaFree(b.p);
    b.p = nullptr;
  }
  b.bytes = bytes;
  if(bytes == 0)
    return OIT_OK;
  CUDA_TRY(c, cudaMalloc(&b.p, bytes));
  return OIT_OK;
}
void devFree(DevBuf& b)
{
  if(b.p)
    cudaFree(b.p);
  b.p     = nullptr;
  b.bytes = 0;
}

void destroyGraphs(OitCtx* c)
{
  for(int set = 0; set < 2; set++)
  {
    for(int half = 0; half < 2; half++)
    {
      if(c->graphExec[set][half])
        cudaGraphExecDestroy(c->graphExec[set][half]);
      if(c->graph[set][half])
        cudaGraphDestroy(c->graph[set][half]);
      c->graphExec[set][half] = nullptr;
      c->graph[set][half]     = nullptr;
    }
    c->graphBuilt[set] = false;
  }
}

void freeBins(BinBuffers& b)
{
  cudaFree(b.lb);
  cudaFree(b.pairKey[0]);
  cudaFree(b.pairKey[1]);
  cudaFree(b.pairVal[0]);
  cudaFree(b.pairVal[1]);
  cudaFree(b.tileStart);
  cudaFree(b.clipEntries);
  cudaFree(b.scratch);
  cudaFree(b.tileOrder);
  b = BinBuffers{};
}

// (re)allocates the binning buffers of one draw for `triCount` triangles and `pairCapacity` pairs
int allocBins(OitCtx* c, BinBuffers& b, size_t triCount, size_t pairCapacity, size_t clipCapacity = 0)
{
  // pieces of near-clipped triangles (oit_clip.cuh): rare, so the table starts small and grows like the pair buffers
  clipCapacity = std::max<size_t>(std::max<size_t>(clipCapacity, b.clipCapacity), 4096);
  const size_t numTiles = (size_t)c->fp.tilesX * c->fp.tileRowsLocal;
  freeBins(b);
  b.pairCapacity = pairCapacity;
  b.triCapacity  = triCount;
  b.scratchWords = binScratchWords(triCount, pairCapacity, numTiles);
  b.lbBytes = binLookbackBytes(triCount, pairCapacity);
  CUDA_TRY(c, cudaMalloc(&b.lb, b.lbBytes));
  CUDA_TRY(c, cudaMemset(b.lb, 0, b.lbBytes));
  b.pairInfo = b.lb;
  for(int i = 0; i < 2; i++)
  {
    CUDA_TRY(c, cudaMalloc(&b.pairKey[i], std::max<size_t>(pairCapacity, 1) * sizeof(uint32_t)));
    CUDA_TRY(c, cudaMalloc(&b.pairVal[i], std::max<size_t>(pairCapacity, 1) * sizeof(uint32_t)));
  }
  CUDA_TRY(c, cudaMalloc(&b.tileStart, (numTiles + 1) * sizeof(uint32_t)));
  b.clipCapacity = clipCapacity;
  CUDA_TRY(c, cudaMalloc(&b.clipEntries, clipCapacity * sizeof(ClipEntry)));
  CUDA_TRY(c, cudaMalloc(&b.tileOrder, std::max<size_t>(numTiles, 1) * sizeof(uint32_t)));
  CUDA_TRY(c, cudaMalloc(&b.scratch, b.scratchWords * sizeof(uint32_t)));
  c->graphValid = false;  // the captured frame refers to the old buffers
  return OIT_OK;
}

void splitObjects(const OitCtx* c, uint32_t& numTransparent, uint32_t& numOpaque)
{
  // oitRender.cpp:68-78
  const int numObjects = c->idxPerObj ? (int)(c->nIndices / c->idxPerObj) : 0;
  int       nt         = (int)(((long long)numObjects * c->cfg.percentTransparent) / 100);
  nt                   = std::min(std::max(nt, 0), numObjects);
  numTransparent       = (uint32_t)nt;
  numOpaque            = (uint32_t)(numObjects - nt);
}

void record(OitCtx* c, int id, cudaStream_t stream = nullptr)
{
  // inside a stream capture the stage events become EXTERNAL event-record nodes, so that they are re-recorded by every
  // replay of the frame graph and cudaEventElapsedTime keeps working on them
  if(!stream)
    stream = c->stream;
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  cudaStreamIsCapturing(stream, &st);
  if(st == cudaStreamCaptureStatusActive)
    cudaEventRecordWithFlags(c->ev[id], stream, cudaEventRecordExternal);
  else
    cudaEventRecord(c->ev[id], stream);
  c->evRecorded[id] = true;
}

// points the frame parameters at buffer set `par` (see OitCtx::stats)
void selectSet(OitCtx* c, int par)
{
  c->par      = par;
  c->fp.stats = (unsigned long long*)c->stats[par].p;
  c->fp.tv      = (TVert*)c->tv[par].p;
  c->fp.tvAttr  = c->tv[par].p ? (float4*)((TVert*)c->tv[par].p + c->nVerts) : nullptr;
  c->fp.ubo   = (const DeviceUbo*)c->uboDev[par].p;
}
int numSets(const OitCtx* c) { return c->pipelined ? 2 : 1; }

DevBuf* bufferOf(OitCtx* c, OitBuffer which)
{
  switch(which)
  {
    case OIT_BUF_ABUFFER: return &c->abuf;
    case OIT_BUF_AUX: return &c->aux;
    case OIT_BUF_AUXSPIN: return &c->spin;
    case OIT_BUF_AUXDEPTH: return &c->adepth;
    case OIT_BUF_COUNTER: return &c->counter;
    case OIT_BUF_COLOR: return &c->color;
    case OIT_BUF_DEPTH: return &c->depth;
    case OIT_BUF_WACCUM: return &c->wacc;
    case OIT_BUF_WREVEAL: return &c->wrev;
    case OIT_BUF_FINAL: return &c->fin;
    case OIT_BUF_FRAME: return &c->frame;
  }
  return nullptr;
}

// bins one draw range, asynchronously: the pair count stays on the device (BinBuffers::pairInfo)
int binDraw(OitCtx* c, int which, uint32_t firstObj, uint32_t numObj, bool cullBack, cudaStream_t stream)
{
  BinBuffers&    b         = c->bins[c->par][which];
  const uint32_t triPerObj = c->idxPerObj / 3;
  const uint32_t firstTri = firstObj * triPerObj, triCount = numObj * triPerObj;
  c->drawTris[which]      = triCount;
  if(triCount == 0)
    return OIT_OK;
  c->fp.clipEntries  = b.clipEntries;  // k_bin_emit fills the draw's clip table
  c->fp.clipCapacity = (uint32_t)b.clipCapacity;
  c->launches += launchBin(c->fp, b, firstTri, triCount, cullBack, &c->sortedBuf[c->par][which], stream);
  return OIT_OK;
}

// after a frame: did a pair buffer overflow?  If so grow it (the frame has to be rendered again).
int growBinsIfNeeded(OitCtx* c, bool* grown)
{
  *grown = false;
  // oit_render mirrors the counters into pinned host memory at the end of the (captured) frame: no extra round trips
  const bool         mirrored = c->mirrorValid;
  unsigned long long overflow = 0;
  if(mirrored)
    overflow = c->hostMirror[STAT_OVERFLOW];
  else
    CUDA_TRY(c, cudaMemcpy(&overflow, (unsigned long long*)c->stats[c->par].p + STAT_OVERFLOW, sizeof(overflow), cudaMemcpyDeviceToHost));
  for(int which = 0; which < 2; which++)
  {
    BinBuffers& b = c->bins[c->par][which];
    if(!b.pairInfo || c->drawTris[which] == 0)
      continue;
    uint32_t info[4] = {0, 0, 0, 0};  // pairs present, pairs wanted, clip entries wanted
    if(mirrored)
      memcpy(info, c->hostMirror + NUM_STAT_SLOTS + 2 * which, sizeof(info));
    else
      CUDA_TRY(c, cudaMemcpy(info, b.pairInfo, sizeof(info), cudaMemcpyDeviceToHost));
    c->pairTotal[which] = info[1];
    if(overflow && (info[1] > b.pairCapacity || info[2] > b.clipCapacity))
    {
      const size_t tris  = b.triCapacity;
      const size_t pairs = info[1] > b.pairCapacity ? (size_t)info[1] + info[1] / 4 + 1024 : b.pairCapacity;
      const size_t clips = info[2] > b.clipCapacity ? (size_t)info[2] + info[2] / 4 + 1024 : b.clipCapacity;
      // both sets see the same scene: they grow together
      for(int set = 0; set < numSets(c); set++)
      {
        const int r = allocBins(c, c->bins[set][which], tris, pairs, clips);
        if(r != OIT_OK)
          return r;
      }
      *grown = true;
    }
  }
  return OIT_OK;
}

void useBins(OitCtx* c, int which)
{
  const BinBuffers& b = c->bins[c->par][which];
  c->fp.pairTri       = b.pairVal[c->sortedBuf[c->par][which]];
  c->fp.tileStart     = b.tileStart;
  c->fp.tileOrder     = b.tileOrder;
  c->fp.clipEntries   = b.clipEntries;
  c->fp.clipCapacity  = (uint32_t)b.clipCapacity;
}

int ensureSceneBins(OitCtx* c)
{
  uint32_t nt, no;
  splitObjects(c, nt, no);
  const size_t triPerObj = c->idxPerObj / 3;
  const size_t need[2]   = {nt * triPerObj, no * triPerObj};
  for(int set = 0; set < numSets(c); set++)
    for(int i = 0; i < 2; i++)
    {
      if(need[i] == 0)
        continue;
      BinBuffers& b = c->bins[set][i];
      if(b.lb == nullptr || b.triCapacity < need[i])
      {
        // (a set that is created late starts with the capacity its twin has already grown to)
        const size_t cap = std::max<size_t>(std::max<size_t>(need[i] * 2, 1u << 16), c->bins[set ^ 1][i].pairCapacity);
        const int    r   = allocBins(c, b, need[i], cap, c->bins[set ^ 1][i].clipCapacity);
        if(r != OIT_OK)
          return r;
      }
    }
  return OIT_OK;
}

}  // namespace

// ====================================================================================================================
extern "C" {

static int finishFrame(OitCtx* c);

int oit_abi_version(void) { return OIT_B200_ABI_VERSION; }

// Host-only proof that the frame kernels' sRGB encoder (oit_device.cuh: enc8 -- the code of the value's bucket plus one
// if the value has passed the bucket's one threshold) equals the definition (the largest k with c >= thr[k]) for EVERY
// float in [0, 1] (stride 1: all 1,065,353,217 of them) and for the values outside, negative zero and NaN.
int oit_selfcheck_srgb_encoder(uint32_t stride, uint64_t* checked, uint64_t* mismatches)
{
  if(stride == 0 || !checked || !mismatches)
    return OIT_ERR_INVALID_ARG;
  alignas(16) static unsigned char tableBytes[SRGB_TABLE_BYTES];
  if(!buildTables(tableBytes))
    return OIT_ERR_CUDA;
  const float*         t      = reinterpret_cast<const float*>(tableBytes);
  const unsigned char* bucket = tableBytes + SRGB_TABLE_FLOATS * 4;
  auto                 device = [&](float c) -> uint32_t {  // enc8 of oit_device.cuh, statement for statement
    const float cc = fminf(fmaxf(c, 0.f), 1.f);
    uint32_t    bits;
    memcpy(&bits, &cc, 4);
    const uint32_t idx = std::max(bits >> 16, SRGB_BUCKET_BASE) - SRGB_BUCKET_BASE;
    const uint32_t k   = bucket[idx];
    return k + (cc >= t[TAB_THR + k + 1] ? 1u : 0u);
  };
  const uint32_t oneBits = 0x3F800000u;
  const unsigned nThreads = std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  std::vector<uint64_t>    bad(nThreads, 0), cnt(nThreads, 0);
  std::vector<std::thread> pool;
  for(unsigned w = 0; w < nThreads; w++)
    pool.emplace_back([&, w]() {
      const uint64_t span = ((uint64_t)oneBits + 1 + nThreads - 1) / nThreads;
      uint64_t       b0 = w * span, b1 = std::min<uint64_t>((uint64_t)oneBits + 1, b0 + span);
      b0 += (stride - b0 % stride) % stride;
      for(uint64_t b = b0; b < b1; b += stride)
      {
        const uint32_t bits = (uint32_t)b;
        float          c;
        memcpy(&c, &bits, 4);
        bad[w] += device(c) != hostEnc8(t, c);
        cnt[w]++;
      }
    });
  for(auto& th : pool)
    th.join();
  uint64_t nBad = 0, n = 0;
  for(unsigned w = 0; w < nThreads; w++)
  {
    nBad += bad[w];
    n += cnt[w];
  }
  const float specials[] = {-0.f, -1.f, 1.0000001f, 2.f, 1e30f, -1e30f, std::numeric_limits<float>::infinity(), -std::numeric_limits<float>::infinity(),
                            std::numeric_limits<float>::quiet_NaN(), 1e-45f, 1.17549435e-38f};
  for(float c : specials)
  {
    // the definition on a clamped value (NaN -> 0, as fminf(fmaxf(c, 0), 1) does)
    const float cc = c != c ? 0.f : std::min(std::max(c, 0.f), 1.f);
    nBad += device(c) != hostEnc8(t, cc);
    n++;
  }
  *checked    = n;
  *mismatches = nBad;
  return OIT_OK;
}

void oit_default_config(OitConfig* cfg)
{
  memset(cfg, 0, sizeof(*cfg));
  cfg->algorithm                     = OIT_SPINLOCK;  // oit.h:66-76
  cfg->oitLayers                     = 8;
  cfg->linkedListAllocatedPerElement = 10;
  cfg->percentTransparent            = 100;
  cfg->tailBlend                     = 1;
  cfg->interlockIsOrdered            = 1;
  cfg->numObjects                    = 1024;
  cfg->subdiv                        = 16;
  cfg->scaleMin                      = 0.1f;
  cfg->scaleWidth                    = 0.9f;
  cfg->aaType                        = OIT_AA_NONE;
  cfg->width                         = 1280;
  cfg->height                        = 720;
  cfg->device                        = 0;
  cfg->bandCount                     = 1;
  cfg->bandIndex                     = 0;
  cfg->stripRows                     = 32;
}

const char* oit_last_error(const OitCtx* ctx) { return ctx ? ctx->error.c_str() : g_createError.c_str(); }

int oit_create(const OitConfig* cfg, OitCtx** out)
{
  if(!cfg || !out)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if(cfg->algorithm >= OIT_NUM_ALGORITHMS)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "unknown algorithm");
  if(cfg->aaType >= OIT_NUM_AATYPES)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "unknown aaType");
  if(cfg->oitLayers < 1 || cfg->oitLayers > 32)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "oitLayers must be in 1..32");
  if(cfg->width == 0 || cfg->height == 0 || cfg->width > 16384 || cfg->height > 16384)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "bad target size");
  if(cfg->linkedListAllocatedPerElement < 1 && cfg->algorithm == OIT_LINKEDLIST)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "linkedListAllocatedPerElement must be >= 1");
  if(cfg->percentTransparent < 0 || cfg->percentTransparent > 100)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "percentTransparent must be in 0..100");
  const uint32_t bandCount = cfg->bandCount ? cfg->bandCount : 1;
  if(cfg->bandIndex >= bandCount)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "bandIndex >= bandCount");
  const uint32_t stripRows = cfg->stripRows ? cfg->stripRows : 32;
  if(stripRows % TILE_H)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "stripRows must be a multiple of 16");

  int nDev = 0;
  if(cudaGetDeviceCount(&nDev) != cudaSuccess || nDev == 0)
    return fail(nullptr, OIT_ERR_CUDA, "no CUDA device: liboit_b200 has no CPU fallback");
  if(cfg->device < 0 || cfg->device >= nDev)
    return fail(nullptr, OIT_ERR_INVALID_ARG, "bad device ordinal");

  OitCtx* c = new(std::nothrow) OitCtx();
  if(!c)
    return fail(nullptr, OIT_ERR_OUT_OF_MEMORY, "host allocation failed");
  c->cfg           = *cfg;
  c->cfg.bandCount = bandCount;
  c->cfg.stripRows = stripRows;
  c->stripRows     = stripRows;
  // State::recomputeAntialiasingSettings (oit.h:88-115)
  switch(cfg->aaType)
  {
    case OIT_AA_MSAA_4X: c->msaa = 4; break;
    case OIT_AA_SSAA_4X: c->msaa = 4; c->sampleShading = true; break;
    case OIT_AA_SUPER_4X: c->supersample = 2; break;
    case OIT_AA_MSAA_8X: c->msaa = 8; break;
    case OIT_AA_SSAA_8X: c->msaa = 8; c->sampleShading = true; break;
    default: break;
  }
  c->coverage = c->msaa > 1 && !c->sampleShading;
  c->bufW     = cfg->width * c->supersample;
  c->bufH     = cfg->height * c->supersample;

  auto cleanupFail = [&](int code) {
    g_createError = c->error;
    oit_destroy(c);
    return code;
  };
#define CREATE_TRY(call)                                                                                                         \
  do                                                                                                                             \
  {                                                                                                                              \
    int r__ = (call);                                                                                                            \
    if(r__ != OIT_OK)                                                                                                            \
      return cleanupFail(r__);                                                                                                   \
  } while(0)
#define CREATE_CUDA(call)                                                                                                        \
  do                                                                                                                             \
  {                                                                                                                              \
    cudaError_t e__ = (call);                                                                                                    \
    if(e__ != cudaSuccess)                                                                                                       \
    {                                                                                                                            \
      c->error = std::string(#call) + ": " + cudaGetErrorString(e__);                                                            \
      return cleanupFail(e__ == cudaErrorMemoryAllocation ? OIT_ERR_OUT_OF_MEMORY : OIT_ERR_CUDA);                                \
    }                                                                                                                            \
  } while(0)

  CREATE_CUDA(cudaSetDevice(cfg->device));
  {
    // the geometry stream gets the higher priority: its small kernels must find their way between the raster CTAs
    int prLeast = 0, prGreatest = 0;
    CREATE_CUDA(cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
    CREATE_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prLeast));
    CREATE_CUDA(cudaStreamCreateWithPriority(&c->geoStream, cudaStreamNonBlocking, prGreatest));
    for(int i = 0; i < 2; i++)
    {
      CREATE_CUDA(cudaEventCreateWithFlags(&c->evGeoDone[i], cudaEventDisableTiming));
      CREATE_CUDA(cudaEventCreateWithFlags(&c->evRasterDone[i], cudaEventDisableTiming));
    }
    CREATE_CUDA(cudaEventCreateWithFlags(&c->evMain, cudaEventDisableTiming));
    c->pipelined = getenv("OIT_B200_NO_PIPELINE") == nullptr;
  }
  for(int i = 0; i < NUM_EVENTS; i++)
    CREATE_CUDA(cudaEventCreate(&c->ev[i]));
  CREATE_CUDA(cudaMallocHost(&c->hostScalar, 64));
  CREATE_CUDA(cudaMallocHost(&c->hostUbo, UBO_RING * sizeof(DeviceUbo)));
  for(int i = 0; i < UBO_RING; i++)
    CREATE_CUDA(cudaEventCreateWithFlags(&c->uboEv[i], cudaEventDisableTiming));
  c->asyncRender = getenv("OIT_B200_SYNC_RENDER") == nullptr;
  CREATE_CUDA(cudaMallocHost(&c->hostMirror, (NUM_STAT_SLOTS + 4) * sizeof(unsigned long long)));
  memset(c->hostMirror, 0, (NUM_STAT_SLOTS + 4) * sizeof(unsigned long long));
  memset(c->hostUbo, 0, UBO_RING * sizeof(DeviceUbo));
  c->useGraph = getenv("OIT_B200_NO_GRAPH") == nullptr;

  // tile / band geometry
  FrameParams& fp    = c->fp;
  fp.W               = (int)c->bufW;
  fp.H               = (int)c->bufH;
  fp.msaa            = c->msaa;
  fp.sampleShading   = c->sampleShading ? 1 : 0;
  fp.coverage        = c->coverage ? 1 : 0;
  fp.L               = (int)cfg->oitLayers;
  fp.tailBlend       = cfg->tailBlend ? 1 : 0;
  fp.layers          = c->sampleShading ? c->msaa : 1;
  fp.algorithm       = (int)cfg->algorithm;
  fp.supersample     = c->supersample;
  fp.fused           = 0;
  // oit_render fuses composite + resolve into the colour-pass kernel (tile colour in shared memory) unless the caller
  // asked to keep the intermediate m_colorImage (reserved[0] bit 0) or OIT_LAYERS exceeds the fused kernel's arrays
  c->fuseFrame = (cfg->reserved[0] & 1u) == 0 && cfg->oitLayers <= 8 && getenv("OIT_B200_NO_FUSE") == nullptr;
  fp.tilesX          = (fp.W + TILE_W - 1) / TILE_W;
  fp.tileRowsGlobal  = (fp.H + TILE_H - 1) / TILE_H;
  fp.stripTileRows   = (int)(stripRows * c->supersample) / TILE_H;
  fp.bandCount       = (int)bandCount;
  fp.bandIndex       = (int)cfg->bandIndex;
  fp.tileRowsLocal   = 0;
  uint32_t localBufH = 0;
  for(int R = 0; R < fp.tileRowsGlobal; R++)
    if(tileRowOwner(R, fp.stripTileRows, fp.bandCount) == fp.bandIndex)
    {
      fp.tileRowsLocal++;
      localBufH += (uint32_t)std::min(TILE_H, fp.H - R * TILE_H);
    }
  {
    // which global tile rows this band owns, and where they sit in its buffers (the binning looks rows up here)
    std::vector<int32_t> rowLocal((size_t)std::max(fp.tileRowsGlobal, 1), -1);
    for(int R = 0; R < fp.tileRowsGlobal; R++)
      if(tileRowOwner(R, fp.stripTileRows, fp.bandCount) == fp.bandIndex)
        rowLocal[R] = tileRowToLocal(R, fp.stripTileRows, fp.bandCount);
    CREATE_TRY(devAlloc(c, c->rowLocal, rowLocal.size() * sizeof(int32_t)));
    CREATE_CUDA(cudaMemcpy(c->rowLocal.p, rowLocal.data(), rowLocal.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    fp.rowLocal = (const int32_t*)c->rowLocal.p;
  }
  c->localBufH = localBufH;
  c->localOutH = localBufH / c->supersample;
  fp.localH    = (int)localBufH;

  // createFrameImages (oit.cpp:84-163), sized for the rows this band owns
  const size_t P     = (size_t)c->bufW * localBufH;
  size_t       words = 0;
  fp.capacity        = 0;
  switch(cfg->algorithm)
  {
    case OIT_SIMPLE:
    case OIT_SPINLOCK:
    case OIT_INTERLOCK: words = P * cfg->oitLayers * (c->coverage ? 4 : 2); break;
    case OIT_LINKEDLIST: {
      words = P * (size_t)cfg->linkedListAllocatedPerElement * 4;
      // uint32 arithmetic like the reference (oit.cpp:125,151); node indices must stay below 2^31 (oitLinkedList.frag.glsl:82)
      const unsigned long long cap = (unsigned long long)cfg->linkedListAllocatedPerElement * P * fp.layers;
      if(cap >= (1ull << 31))
      {
        c->error = "linked-list pool would exceed 2^31 nodes";
        return cleanupFail(OIT_ERR_INVALID_ARG);
      }
      fp.capacity = (uint32_t)cap;
      break;
    }
    case OIT_LOOP: words = P * cfg->oitLayers * 2; break;
    case OIT_LOOP64: words = P * cfg->oitLayers * 2; break;
    default: break;
  }
  words *= fp.layers;
  // the linked-list pool has at least node 0 + one usable node so that the index arithmetic is always in bounds
  CREATE_TRY(devAlloc(c, c->abuf, std::max<size_t>(words, 4) * 4));
  if(cfg->algorithm != OIT_WEIGHTED)
    CREATE_TRY(devAlloc(c, c->aux, P * fp.layers * 4));
  if(cfg->algorithm == OIT_SPINLOCK)
    CREATE_TRY(devAlloc(c, c->spin, P * fp.layers * 4));
  if(cfg->algorithm == OIT_SPINLOCK || cfg->algorithm == OIT_INTERLOCK)
    CREATE_TRY(devAlloc(c, c->adepth, P * fp.layers * 4));
  if(cfg->algorithm == OIT_LINKEDLIST)
    CREATE_TRY(devAlloc(c, c->counter, 4));
  if(cfg->algorithm == OIT_WEIGHTED)
  {
    CREATE_TRY(devAlloc(c, c->wacc, P * c->msaa * 8));
    CREATE_TRY(devAlloc(c, c->wrev, ((P * c->msaa + 1) / 2) * 4));
  }
  CREATE_TRY(devAlloc(c, c->color, P * c->msaa * 4));
  if(cfg->percentTransparent < 100)
    CREATE_TRY(devAlloc(c, c->depth, P * c->msaa * 4));
  CREATE_TRY(devAlloc(c, c->fin, std::max<size_t>((size_t)cfg->width * c->localOutH, 1) * 4));
  CREATE_TRY(devAlloc(c, c->tables, SRGB_TABLE_BYTES));
  for(int set = 0; set < numSets(c); set++)
  {
    CREATE_TRY(devAlloc(c, c->stats[set], NUM_STAT_SLOTS * sizeof(unsigned long long)));
    CREATE_CUDA(cudaMemset(c->stats[set].p, 0, c->stats[set].bytes));
    CREATE_TRY(devAlloc(c, c->uboDev[set], sizeof(DeviceUbo)));
  }
  alignas(16) unsigned char tableBytes[SRGB_TABLE_BYTES];
  const float*              tables = reinterpret_cast<const float*>(tableBytes);
  if(!buildTables(tableBytes))
  {
    c->error = "internal error: the sRGB encoder's bucket table is not one-threshold-per-bucket";
    return cleanupFail(OIT_ERR_CUDA);
  }
  CREATE_CUDA(cudaMemcpy(c->tables.p, tableBytes, SRGB_TABLE_BYTES, cudaMemcpyHostToDevice));
  // clear colour (0.2, 0.2, 0.2, 0.2) linear -> B8G8R8A8_SRGB (oitRender.cpp:90)
  const uint32_t rgb = hostEnc8(tables, 0.2f);
  fp.clearColor      = rgb | (rgb << 8) | (rgb << 16) | ((uint32_t)rintf(0.2f * 255.0f) << 24);

  fp.abuf    = (uint32_t*)c->abuf.p;
  fp.aux     = (uint32_t*)c->aux.p;
  fp.spin    = (uint32_t*)c->spin.p;
  fp.adepth  = (uint32_t*)c->adepth.p;
  fp.counter = (uint32_t*)c->counter.p;
  fp.color   = (uint32_t*)c->color.p;
  fp.depth   = (float*)c->depth.p;
  fp.wacc    = (uint16_t*)c->wacc.p;
  fp.wrev    = (uint16_t*)c->wrev.p;
  fp.fin     = (uint32_t*)c->fin.p;
  fp.tables  = (const float*)c->tables.p;
  selectSet(c, 0);
  *out          = c;
  return OIT_OK;
#undef CREATE_TRY
#undef CREATE_CUDA
}

int oit_destroy(OitCtx* c)
{
  if(!c)
    return OIT_OK;
  cudaSetDevice(c->cfg.device);
  if(c->geoStream)
    cudaStreamSynchronize(c->geoStream);
  if(c->stream)
    cudaStreamSynchronize(c->stream);
  // the captured frame graphs reference the NCCL communicator: they have to go first
  destroyGraphs(c);
  gatherDestroy(c->gather);
  if(c->peers)
  {
    peerDestroy(c->peers);
    c->frame = DevBuf{};  // part of the peer chunk
  }
  if(!c->finOwned)
    c->fin = DevBuf{};  // a slice of gatherBuf
  for(DevBuf* b : {&c->abuf, &c->aux, &c->spin, &c->adepth, &c->counter, &c->color, &c->depth, &c->wacc, &c->wrev, &c->fin,
                   &c->tables, &c->rowLocal, &c->pushQueue, &c->stats[0], &c->stats[1], &c->tv[0], &c->tv[1], &c->uboDev[0], &c->uboDev[1], &c->gatherBuf, &c->frame,
                   &c->sphTable, &c->sphUnitPos, &c->sphUnitTri})
    devFree(*b);
  if(c->sceneOwned)
  {
    devFree(c->verts);
    devFree(c->indices);
  }
  for(int set = 0; set < 2; set++)
    for(int which = 0; which < 2; which++)
      freeBins(c->bins[set][which]);
  for(int i = 0; i < 2; i++)
  {
    if(c->evGeoDone[i])
      cudaEventDestroy(c->evGeoDone[i]);
    if(c->evRasterDone[i])
      cudaEventDestroy(c->evRasterDone[i]);
  }
  if(c->evMain)
    cudaEventDestroy(c->evMain);
  for(int i = 0; i < NUM_EVENTS; i++)
    if(c->ev[i])
      cudaEventDestroy(c->ev[i]);
  if(c->hostScalar)
    cudaFreeHost(c->hostScalar);
  if(c->hostUbo)
    cudaFreeHost(c->hostUbo);
  for(int i = 0; i < UBO_RING; i++)
    if(c->uboEv[i])
      cudaEventDestroy(c->uboEv[i]);
  if(c->hostMirror)
    cudaFreeHost(c->hostMirror);
  if(c->geoStream)
    cudaStreamDestroy(c->geoStream);
  if(c->stream)
    cudaStreamDestroy(c->stream);
  delete c;
  return OIT_OK;
}

int oit_get_config(const OitCtx* c, OitConfig* out)
{
  if(!c || !out)
    return OIT_ERR_INVALID_ARG;
  *out = c->cfg;
  return OIT_OK;
}

int oit_get_dims(const OitCtx* c, uint32_t* bufW, uint32_t* bufH, uint32_t* msaa, uint32_t* sampleShading, uint32_t* localRows)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  if(bufW)
    *bufW = c->bufW;
  if(bufH)
    *bufH = c->bufH;
  if(msaa)
    *msaa = (uint32_t)c->msaa;
  if(sampleShading)
    *sampleShading = c->sampleShading ? 1u : 0u;
  if(localRows)
    *localRows = c->localOutH;
  return OIT_OK;
}

static int installScene(OitCtx* c, uint32_t nVerts, uint32_t nIndices, uint32_t indicesPerObject)
{
  c->nVerts     = nVerts;
  c->nIndices   = nIndices;
  c->idxPerObj  = indicesPerObject;
  c->graphValid = false;
  for(int set = 0; set < numSets(c); set++)
  {
    const int r = devAlloc(c, c->tv[set], (size_t)nVerts * (sizeof(TVert) + ATTR_FLOATS * sizeof(float)));  // [TVert table][attribute records]
    if(r != OIT_OK)
      return r;
  }
  c->fp.verts   = (const float*)c->verts.p;
  c->fp.indices = (const uint32_t*)c->indices.p;
  c->fp.nVerts  = nVerts;
  selectSet(c, c->par);
  return ensureSceneBins(c);
}

// every index must refer to an existing vertex: checked on the device, where the index buffer already is (a host loop over
// the sample's 3.1 M indices costs more than the frame).  On failure the context is left without a scene.
static int validateIndices(OitCtx* c, const uint32_t* dIndices, uint32_t nIndices, uint32_t nVerts)
{
  unsigned long long* flag = (unsigned long long*)c->stats[0].p + STAT_SCRATCH;  // no frame is in flight here
  CUDA_TRY(c, cudaMemsetAsync(flag, 0, sizeof(*flag), c->stream));
  launchValidateIndices(dIndices, nIndices, nVerts, flag, c->stream);
  unsigned long long bad = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&bad, flag, sizeof(bad), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if(!bad)
    return OIT_OK;
  if(c->sceneOwned)
  {
    devFree(c->verts);
    devFree(c->indices);
  }
  c->verts = c->indices = DevBuf{};
  c->sceneOwned         = true;
  c->nVerts = c->nIndices = 0;
  c->graphValid           = false;
  return fail(c, OIT_ERR_INVALID_ARG, "index out of range");
}

int oit_set_scene(OitCtx* c, const void* vertices, uint32_t nVerts, const uint32_t* indices, uint32_t nIndices, uint32_t indicesPerObject)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  if(!vertices || !indices || nVerts == 0 || indicesPerObject == 0 || indicesPerObject % 3 || nIndices % indicesPerObject)
    return fail(c, OIT_ERR_INVALID_ARG, "bad scene arguments");
  c->mainDirty = true;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  if(!c->sceneOwned)
  {
    c->verts   = DevBuf{};
    c->indices = DevBuf{};
  }
  c->sceneOwned = true;
  if(c->verts.bytes != (size_t)nVerts * 40)
  {
    c->graphValid = false;
    int r         = devAlloc(c, c->verts, (size_t)nVerts * 40);
    if(r != OIT_OK)
      return r;
  }
  if(c->indices.bytes != (size_t)nIndices * 4)
  {
    c->graphValid = false;
    int r         = devAlloc(c, c->indices, (size_t)nIndices * 4);
    if(r != OIT_OK)
      return r;
  }
  CUDA_TRY(c, cudaMemcpyAsync(c->verts.p, vertices, (size_t)nVerts * 40, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(c->indices.p, indices, (size_t)nIndices * 4, cudaMemcpyHostToDevice, c->stream));
  int r = validateIndices(c, (const uint32_t*)c->indices.p, nIndices, nVerts);
  if(r != OIT_OK)
    return r;
  if(c->nVerts != nVerts || c->nIndices != nIndices || c->idxPerObj != indicesPerObject || !c->tv[0].p)
    return installScene(c, nVerts, nIndices, indicesPerObject);
  c->fp.verts   = (const float*)c->verts.p;
  c->fp.indices = (const uint32_t*)c->indices.p;
  return OIT_OK;
}

int oit_set_scene_device(OitCtx* c, const void* dVertices, uint32_t nVerts, const uint32_t* dIndices, uint32_t nIndices, uint32_t indicesPerObject)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  if(!dVertices || !dIndices || nVerts == 0 || indicesPerObject == 0 || indicesPerObject % 3 || nIndices % indicesPerObject)
    return fail(c, OIT_ERR_INVALID_ARG, "bad scene arguments");
  c->mainDirty = true;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  if(c->sceneOwned)
  {
    devFree(c->verts);
    devFree(c->indices);
  }
  c->sceneOwned    = false;
  c->verts.p       = const_cast<void*>(dVertices);
  c->verts.bytes   = (size_t)nVerts * 40;
  c->indices.p     = const_cast<uint32_t*>(dIndices);
  c->indices.bytes = (size_t)nIndices * 4;
  const int r      = validateIndices(c, dIndices, nIndices, nVerts);
  if(r != OIT_OK)
    return r;
  return installScene(c, nVerts, nIndices, indicesPerObject);
}

int oit_set_scene_spheres(OitCtx* c, const OitSphere* spheres, uint32_t nSpheres, int32_t subdiv)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  OitConfig sizes   = c->cfg;
  sizes.numObjects  = (int32_t)nSpheres;
  sizes.subdiv      = subdiv;
  uint32_t nVerts = 0, nIndices = 0, ipo = 0;
  if(!spheres || nSpheres == 0 || nSpheres > 0x7FFFFFFFu || oit_scene_sizes(&sizes, &nVerts, &nIndices, &ipo) != OIT_OK)
    return fail(c, OIT_ERR_INVALID_ARG, "bad scene arguments");
  c->mainDirty = true;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  int r;
  if(c->sphSubdiv != subdiv)
  {
    std::vector<float>    pos;
    std::vector<uint32_t> tri;
    unitSphereTemplate(subdiv, pos, tri);
    if((r = devAlloc(c, c->sphUnitPos, pos.size() * sizeof(float))) != OIT_OK || (r = devAlloc(c, c->sphUnitTri, tri.size() * sizeof(uint32_t))) != OIT_OK)
      return r;
    CUDA_TRY(c, cudaMemcpy(c->sphUnitPos.p, pos.data(), pos.size() * sizeof(float), cudaMemcpyHostToDevice));
    CUDA_TRY(c, cudaMemcpy(c->sphUnitTri.p, tri.data(), tri.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    c->sphSubdiv = subdiv;
  }
  if(c->sphTable.bytes != (size_t)nSpheres * sizeof(OitSphere) && (r = devAlloc(c, c->sphTable, (size_t)nSpheres * sizeof(OitSphere))) != OIT_OK)
    return r;
  if(!c->sceneOwned)
  {
    c->verts   = DevBuf{};
    c->indices = DevBuf{};
  }
  c->sceneOwned = true;
  if(c->verts.bytes != (size_t)nVerts * 40)
  {
    c->graphValid = false;
    if((r = devAlloc(c, c->verts, (size_t)nVerts * 40)) != OIT_OK)
      return r;
  }
  if(c->indices.bytes != (size_t)nIndices * 4)
  {
    c->graphValid = false;
    if((r = devAlloc(c, c->indices, (size_t)nIndices * 4)) != OIT_OK)
      return r;
  }
  // the only upload: 32 bytes per object; the flattening runs where the mesh is needed
  CUDA_TRY(c, cudaMemcpyAsync(c->sphTable.p, spheres, (size_t)nSpheres * sizeof(OitSphere), cudaMemcpyHostToDevice, c->stream));
  launchExpandSpheres((const float*)c->sphTable.p, nSpheres, (const float*)c->sphUnitPos.p, nVerts / nSpheres, (const uint32_t*)c->sphUnitTri.p,
                      ipo, (float*)c->verts.p, (uint32_t*)c->indices.p, c->stream);
  CUDA_TRY(c, cudaGetLastError());
  if(c->nVerts != nVerts || c->nIndices != nIndices || c->idxPerObj != ipo || !c->tv[0].p)
    return installScene(c, nVerts, nIndices, ipo);
  c->fp.verts   = (const float*)c->verts.p;
  c->fp.indices = (const uint32_t*)c->indices.p;
  return OIT_OK;
}

// uploads the per-frame UBO into buffer set c->par, ordered on `stream`
static int uploadSceneData(OitCtx* c, const OitSceneData* ubo, cudaStream_t stream)
{
  c->ubo = *ubo;
  // updateUniformBuffer (main.cpp:628-637) + createFrameImages (oit.cpp:102-152) own these fields
  c->ubo.viewport[0] = (int32_t)c->bufW;
  c->ubo.viewport[1] = (int32_t)c->bufH;
  c->ubo.viewport[2] = (int32_t)(c->bufW * c->bufH);
  c->ubo.linkedListAllocatedPerElement =
      c->cfg.algorithm == OIT_LINKEDLIST ? c->fp.capacity : c->cfg.oitLayers * (uint32_t)c->fp.layers;
  // frames may be in flight: the upload goes through a ring of pinned staging slots; waiting for a slot's previous upload
  // also bounds the number of frames the host can run ahead of the device
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  c->uboSlot      = (c->uboSlot + 1) % UBO_RING;
  DeviceUbo* slot = c->hostUbo + c->uboSlot;
  CUDA_TRY(c, cudaEventSynchronize(c->uboEv[c->uboSlot]));
  memcpy(slot->projView, ubo->projViewMatrix, sizeof(float) * 16);
  memcpy(slot->view, ubo->viewMatrix, sizeof(float) * 16);
  slot->alphaMin   = ubo->alphaMin;
  slot->alphaWidth = ubo->alphaWidth;
  CUDA_TRY(c, cudaMemcpyAsync(c->uboDev[c->par].p, slot, sizeof(DeviceUbo), cudaMemcpyHostToDevice, stream));
  CUDA_TRY(c, cudaEventRecord(c->uboEv[c->uboSlot], stream));
  c->haveUbo       = true;
  return OIT_OK;
}

int oit_set_scene_data(OitCtx* c, const OitSceneData* ubo)
{
  if(!c || !ubo)
    return OIT_ERR_INVALID_ARG;
  // (on the main stream: ordered behind every frame in flight, whose raster halves follow their geometry halves)
  c->mainDirty = true;
  return uploadSceneData(c, ubo, c->stream);
}

// The geometry half of a frame on `stream`: statistics reset, vertex stage, binning of both draws (into buffer set c->par)
static int issueGeometry(OitCtx* c, cudaStream_t stream)
{
  record(c, EV_START, stream);
  CUDA_TRY(c, cudaMemsetAsync(c->stats[c->par].p, 0, c->stats[c->par].bytes, stream));
  c->launches += launchTransformVertices(c->fp, stream);
  uint32_t nt, no;
  splitObjects(c, nt, no);
  int r;
  if((r = binDraw(c, 0, 0, nt, false, stream)) != OIT_OK)
    return r;
  if((r = binDraw(c, 1, nt, no, true, stream)) != OIT_OK)
    return r;
  record(c, EV_GEOM, stream);
  return OIT_OK;
}

int oit_begin_frame(OitCtx* c)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  if(!c->verts.p || !c->indices.p)
    return fail(c, OIT_ERR_NO_SCENE, "oit_set_scene has not been called");
  if(!c->haveUbo)
    return fail(c, OIT_ERR_INVALID_ARG, "oit_set_scene_data has not been called");
  c->mainDirty = c->mainDirty || !c->capturing;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(!c->capturing && c->framePending)
    if(const int fr = finishFrame(c); fr != OIT_OK)
      return fr;
  int r = OIT_OK;
  if(!c->capturing)
  {
    // stage-by-stage use: everything on the main stream; the pair buffers must be large enough before anything consumes
    // the bins (oit_render checks once per frame instead, after the whole asynchronous frame)
    c->launches = 0;
    for(bool& b : c->evRecorded)
      b = false;
    selectSet(c, c->par);
    if((r = ensureSceneBins(c)) != OIT_OK)
      return r;
    for(int attempt = 0; attempt < 4; attempt++)
    {
      if((r = issueGeometry(c, c->stream)) != OIT_OK)
        return r;
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      bool grown = false;
      if((r = growBinsIfNeeded(c, &grown)) != OIT_OK)
        return r;
      if(!grown)
        break;
    }
  }
  record(c, EV_RASTER_START);
  // clearTransparent* + colour/depth clear
  {
    // the fused linked-list frame kernel starts every list empty by itself (no imgAux clear), and in split-frame mode the
    // first node of the exchange (k_peer_frame_begin) resets the node counter
    const bool leanLL = c->cfg.algorithm == OIT_LINKEDLIST && linkedListFrameStartsEmpty(c->fp);
    c->launches += launchClears(c->fp, (int)c->cfg.algorithm, c->stream, leanLL, leanLL && c->peers && c->peersOpen);
  }
  record(c, EV_CLEAR);
  CUDA_TRY(c, cudaGetLastError());
  return OIT_OK;
}

int oit_draw_opaque(OitCtx* c)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  c->mainDirty = c->mainDirty || !c->capturing;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(c->drawTris[1] > 0 && c->fp.depth)
  {
    useBins(c, 1);
    c->launches += launchRaster(c->fp, PASS_OPAQUE, c->stream);
  }
  record(c, EV_OPAQUE);
  CUDA_TRY(c, cudaGetLastError());
  return OIT_OK;
}

int oit_draw_transparent(OitCtx* c)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  c->mainDirty = c->mainDirty || !c->capturing;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(c->drawTris[0] > 0)
  {
    useBins(c, 0);
    switch(c->cfg.algorithm)
    {
      case OIT_SIMPLE: c->launches += launchRaster(c->fp, PASS_SIMPLE, c->stream); break;
      case OIT_LINKEDLIST: c->launches += launchRaster(c->fp, PASS_LINKEDLIST, c->stream); break;
      case OIT_LOOP:
        c->launches += launchRaster(c->fp, PASS_LOOP_DEPTH, c->stream);  // oitRender.cpp:269-279
        c->launches += launchRaster(c->fp, PASS_LOOP_COLOR, c->stream);
        break;
      case OIT_LOOP64: c->launches += launchRaster(c->fp, PASS_LOOP64, c->stream); break;
      case OIT_SPINLOCK: c->launches += launchRaster(c->fp, PASS_SPINLOCK, c->stream); break;
      case OIT_INTERLOCK: c->launches += launchRaster(c->fp, PASS_INTERLOCK, c->stream); break;
      case OIT_WEIGHTED: c->launches += launchRaster(c->fp, PASS_WEIGHTED, c->stream); break;
    }
  }
  record(c, EV_COLOR);
  CUDA_TRY(c, cudaGetLastError());
  return OIT_OK;
}

int oit_composite(OitCtx* c)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  c->mainDirty = c->mainDirty || !c->capturing;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(!c->fp.fused)
    c->launches += launchComposite(c->fp, (int)c->cfg.algorithm, c->stream);
  if(!c->fp.fused)  // (fused frame: no stage, and an event-record node costs ~2.6 us of every replayed frame)
    record(c, EV_COMPOSITE);
  CUDA_TRY(c, cudaGetLastError());
  return OIT_OK;
}

int oit_resolve(OitCtx* c)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  c->mainDirty = c->mainDirty || !c->capturing;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(!c->fp.fused)
    c->launches += launchResolve(c->fp, c->supersample, (int)c->cfg.width, (int)c->localOutH, c->stream);
  if(!c->fp.fused)
    record(c, EV_RESOLVE);
  CUDA_TRY(c, cudaGetLastError());
  return OIT_OK;
}

int oit_synchronize(OitCtx* c)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  return OIT_OK;
}

// The raster half of one frame, asynchronously on the main stream: clears, opaque pass, colour pass(es), composite, resolve,
// the band exchange and the statistics mirror.  Reads what the geometry half left in buffer set c->par.
static int issueRaster(OitCtx* c)
{
  int                 r;
  const bool          exchange = c->peers && c->peersOpen;
  unsigned long long* stats    = (unsigned long long*)c->stats[c->par].p;
  if((r = oit_begin_frame(c)) != OIT_OK)  // (c->capturing: only the clears)
    return r;
  if((r = oit_draw_opaque(c)) != OIT_OK)
    return r;
  if(exchange)
  {
    // pusher CTAs: the linked-list frame kernel without sample shading / super-sampling (oit_raster_ll.cu)
    const bool leanLL = c->cfg.algorithm == OIT_LINKEDLIST && linkedListFrameStartsEmpty(c->fp);
    const bool push   = leanLL && c->supersample == 1 && c->pushers > 0 && c->pushQueue.p != nullptr;
    c->fp.pushers     = push ? c->pushers : 0;
    c->fp.pushQueue   = (uint32_t*)c->pushQueue.p;
    // the exchange's first node comes as late as possible (the clears and the opaque pass absorb what skew there is): it
    // allows the NEXT frame into this band's other buffer and checks that every band allows this one
    c->launches += peerFrameBegin(c->peers, stats, leanLL ? c->fp.counter : nullptr,
                                  push ? c->fp.pushQueue + c->fp.tilesX * c->fp.tileRowsLocal : nullptr, c->stream);
    record(c, EV_OPAQUE);  // time spent waiting for the other bands is not the colour pass's
  }
  c->fp.peers = (exchange && c->fp.fused) ? peerTable(c->peers) : nullptr;  // the fused kernel stores to every band
  r           = oit_draw_transparent(c);
  c->fp.peers   = nullptr;
  c->fp.pushers = 0;
  if(r != OIT_OK)
    return r;
  if((r = oit_composite(c)) != OIT_OK)
    return r;
  if((r = oit_resolve(c)) != OIT_OK)
    return r;
  if(exchange)
  {
    if(!c->fp.fused)
      c->launches += peerScatterRows(c->peers, (const uint32_t*)c->fin.p, (int)c->cfg.width, (int)c->localOutH, (int)c->stripRows, c->stream);
    // last node: DONE (with this band's overflow flag), the wait for the PREVIOUS frame's DONE, the statistics mirror
    c->launches += peerFrameEnd(c->peers, stats, c->hostMirror, NUM_STAT_SLOTS,
                                (c->bins[c->par][0].pairInfo && c->drawTris[0] > 0) ? c->bins[c->par][0].pairInfo : nullptr,
                                (c->bins[c->par][1].pairInfo && c->drawTris[1] > 0) ? c->bins[c->par][1].pairInfo : nullptr, c->stream);
    CUDA_TRY(c, cudaGetLastError());
    return OIT_OK;
  }
  // split frame: ONE all-gather of the resolved strips over NVLink + the row interleave, still on the same stream
  if(c->gather)
  {
    const int n = gatherLaunch(c->gather, (uint32_t*)c->gatherBuf.p, (uint32_t*)c->frame.p, (int)c->cfg.width, (int)c->cfg.height,
                               (int)c->stripRows, (int)c->padRows, stats, c->stream, c->error);
    if(n < 0)
      return n;
    c->launches += n;
  }
  // mirror the statistics and the pair counts into pinned host memory as the last nodes of the frame
  CUDA_TRY(c, cudaMemcpyAsync(c->hostMirror, stats, NUM_STAT_SLOTS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  for(int which = 0; which < 2; which++)
    if(c->bins[c->par][which].pairInfo && c->drawTris[which] > 0)
      CUDA_TRY(c, cudaMemcpyAsync(c->hostMirror + NUM_STAT_SLOTS + 2 * which, c->bins[c->par][which].pairInfo, 4 * sizeof(uint32_t),
                                  cudaMemcpyDeviceToHost, c->stream));
  return OIT_OK;
}

// One frame, stage by stage, without host synchronisation.  Pipelined: the geometry half goes to the geometry stream and
// the raster half waits for it with an event, so that the NEXT frame's geometry half (other buffer set) overlaps this
// frame's raster half.  Otherwise everything is issued on the main stream.
static int issueFrameStreams(OitCtx* c)
{
  int r;
  c->launches = 0;
  for(bool& b : c->evRecorded)
    b = false;
  if(c->pipelined)
  {
    if((r = issueGeometry(c, c->geoStream)) != OIT_OK)
      return r;
    CUDA_TRY(c, cudaEventRecord(c->evGeoDone[c->par], c->geoStream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->evGeoDone[c->par], 0));
    if((r = issueRaster(c)) != OIT_OK)
      return r;
    CUDA_TRY(c, cudaEventRecord(c->evRasterDone[c->par], c->stream));
    c->rasterRecorded[c->par] = true;
    return OIT_OK;
  }
  if((r = issueGeometry(c, c->stream)) != OIT_OK)
    return r;
  return issueRaster(c);
}

// The same, captured once per buffer set and replayed: the geometry half and the raster half are separate graphs
// (pipelined), or one graph holds the whole frame.
static int buildGraphs(OitCtx* c)
{
  const int set = c->par;
  int       r   = OIT_OK;
  c->launches   = 0;
  for(bool& b : c->evRecorded)
    b = false;
  cudaError_t e = cudaSuccess;
  if(c->pipelined)
  {
    e = cudaStreamBeginCapture(c->geoStream, cudaStreamCaptureModeThreadLocal);
    if(e == cudaSuccess)
    {
      r = issueGeometry(c, c->geoStream);
      e = cudaStreamEndCapture(c->geoStream, &c->graph[set][0]);
      if(r == OIT_OK && e == cudaSuccess)
        e = cudaGraphInstantiate(&c->graphExec[set][0], c->graph[set][0], 0);
    }
    if(r != OIT_OK || e != cudaSuccess)
      return OIT_ERR_CUDA;
  }
  e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
  if(e == cudaSuccess)
  {
    if(!c->pipelined)
      r = issueGeometry(c, c->stream);
    if(r == OIT_OK)
      r = issueRaster(c);
    e = cudaStreamEndCapture(c->stream, &c->graph[set][1]);
    if(r == OIT_OK && e == cudaSuccess)
      e = cudaGraphInstantiate(&c->graphExec[set][1], c->graph[set][1], 0);
  }
  if(r != OIT_OK || e != cudaSuccess)
    return OIT_ERR_CUDA;
  c->graphBuilt[set]       = true;
  c->graphLaunchCount[set] = c->launches;
  return OIT_OK;
}

static int launchGraphs(OitCtx* c)
{
  const int set = c->par;
  c->launches   = c->graphLaunchCount[set];
  if(c->pipelined)
  {
    // (timing experiment OIT_EXPERIMENT_SKIP_GEO=1: after a few frames the geometry half is not launched any more and the
    // raster halves keep reading the tile lists of the static camera -- what would a free geometry stage buy?)
    static const bool skipGeo = getenv("OIT_EXPERIMENT_SKIP_GEO") != nullptr;
    if(!skipGeo || c->experimentFrames++ < 8)
      CUDA_TRY(c, cudaGraphLaunch(c->graphExec[set][0], c->geoStream));
    CUDA_TRY(c, cudaEventRecord(c->evGeoDone[set], c->geoStream));
    CUDA_TRY(c, cudaStreamWaitEvent(c->stream, c->evGeoDone[set], 0));
  }
  CUDA_TRY(c, cudaGraphLaunch(c->graphExec[set][1], c->stream));
  if(c->pipelined)
  {
    CUDA_TRY(c, cudaEventRecord(c->evRasterDone[set], c->stream));
    c->rasterRecorded[set] = true;
  }
  return OIT_OK;
}

// Enqueues one frame (buffer set c->par, whose UBO has been uploaded) without waiting for it.
static int enqueueFrame(OitCtx* c)
{
  int r = OIT_OK;
  for(int tries = 0; tries < 2; tries++)
  {
    if((r = ensureSceneBins(c)) != OIT_OK)
      return r;
    selectSet(c, c->par);
    c->capturing = true;  // the frame is issued without host synchronisation; overflow is checked once at the end
    {
      // the fused kernel is the transparent colour pass: without transparent triangles the staged kernels resolve the frame
      uint32_t nt = 0, no = 0;
      splitObjects(c, nt, no);
      c->fp.fused = (c->fuseFrame && nt > 0) ? 1 : 0;
      // k-buffer techniques without sample shading: the tile's A-buffer slice also stays in shared memory
      const uint32_t words = onChipWords((int)c->cfg.algorithm, (int)c->cfg.oitLayers, c->coverage ? 1 : 0);
      // measured on B200 (4K, no AA): pays off for the atomic-heavy Loop64 and Spinlock protocols; Simple / Interlock lose
      // more from the lower occupancy (3 instead of 5 CTAs per SM) than they gain, unless OIT_B200_ONCHIP_ALL is set
      const bool worthIt = c->cfg.algorithm == OIT_LOOP64 || c->cfg.algorithm == OIT_SPINLOCK || getenv("OIT_B200_ONCHIP_ALL") != nullptr;
      c->fp.onChip = (c->fp.fused && !c->sampleShading && words > 0 && words * 4u <= ON_CHIP_MAX_BYTES && worthIt
                      && getenv("OIT_B200_NO_ONCHIP") == nullptr)
                         ? 1
                         : 0;
    }
    bool retry = false;
    // (with the band gather, the first frame runs un-captured so that NCCL sets up its connections outside a capture)
    if(c->useGraph && (!c->gather || c->gatherWarm))
    {
      // a frame (~30 kernels) is captured once per buffer set and replayed with one or two graph launches
      if(!c->graphValid)
      {
        destroyGraphs(c);
        c->graphValid = true;
      }
      if(!c->graphBuilt[c->par] && buildGraphs(c) != OIT_OK)
      {
        cudaGetLastError();
        destroyGraphs(c);
        c->useGraph = false;  // fall back to plain stream launches
        retry       = true;
      }
      if(!retry)
        r = launchGraphs(c);
    }
    else
      r = issueFrameStreams(c);
    c->capturing = false;
    c->fp.fused  = 0;
    c->fp.onChip = 0;
    if(!retry)
    {
      if(r == OIT_OK && c->peers && c->peersOpen)
        c->peerSeq++;  // (the device counts the same frames: k_peer_frame_end)
      return r;
    }
  }
  return r;
}

// Completes the frame oit_render enqueued: waits for it, and if a (tile, triangle) pair buffer overflowed (normally only
// the first frame after a scene / camera change) grows it and renders the frame again.  Every entry point that hands
// results to the host, or changes what a frame reads, comes through here; without a pending frame it is a stream sync.
static int finishFrame(OitCtx* c)
{
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(!c->framePending)
  {
    CUDA_TRY(c, cudaStreamSynchronize(c->geoStream));
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    return OIT_OK;
  }
  for(int attempt = 0; attempt < 4; attempt++)
  {
    // (every raster half waits for its geometry half: the main stream drains last)
    CUDA_TRY(c, cudaStreamSynchronize(c->geoStream));
    if(c->peers && c->peersOpen)
    {
      // split frame over peer memory: the bands agree on frame boundaries one frame late, so the LATEST frame is completed
      // here -- every band's strips are in this band's buffer, and every band's overflow flag is known
      peerFrameFlush(c->peers, (unsigned long long*)c->stats[c->par].p, c->hostMirror, c->stream);
      c->frame.p = peerFrame(c->peers, c->peerSeq);
    }
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    c->framePending = false;
    bool grown      = false;
    c->mirrorValid  = true;
    int r           = growBinsIfNeeded(c, &grown);
    c->mirrorValid  = false;
    if(r != OIT_OK)
      return r;
    if(c->hostMirror[STAT_INTERNAL] != 0)
      return fail(c, OIT_ERR_CUDA, "internal error: a look-back chain of the binning timed out");
    if(c->hostMirror[STAT_PEER_TIMEOUT] != 0)
      return fail(c, OIT_ERR_CUDA, "split frame: a band did not reach the frame barrier (peer exchange timed out)");
    // Split frame: the decision to render the frame again is COLLECTIVE.  The exchange itself carried every band's overflow
    // flag (STAT_OVERFLOW_ANY), so all bands repeat the frame -- with a full exchange round -- when any of them had to grow a
    // buffer, and no band is left with the strips of an overflowed attempt.  (Bands call the frame-completing entry points
    // in the same order, so they look at the same frame here.)
    bool repeat = grown;
    if(c->gather || (c->peers && c->peersOpen))
    {
      c->gatherWarm = true;
      repeat        = repeat || c->hostMirror[STAT_OVERFLOW_ANY] != 0;
    }
    if(!repeat)
      return OIT_OK;
    if((r = enqueueFrame(c)) != OIT_OK)
      return r;
    c->framePending = true;
  }
  return fail(c, OIT_ERR_OUT_OF_MEMORY, "the (tile, triangle) pair buffers kept overflowing");
}

int oit_render(OitCtx* c, const OitSceneData* ubo)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  int r;
  if(!ubo)
    return OIT_ERR_INVALID_ARG;
  if(!c->verts.p || !c->indices.p)
    return fail(c, OIT_ERR_NO_SCENE, "oit_set_scene has not been called");
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  // a frame in flight that already reported a pair-buffer overflow (the mirror is pinned host memory written by the frame's
  // last nodes): grow the buffers now instead of letting further frames run with truncated triangle lists
  // (not in split-frame mode: there the bands must take the decision at the same frame, which finishFrame guarantees)
  if(c->framePending && !c->gather && !(c->peers && c->peersOpen)
     && reinterpret_cast<volatile unsigned long long*>(c->hostMirror)[STAT_OVERFLOW] != 0)
    if((r = finishFrame(c)) != OIT_OK)
      return r;
  if(c->pipelined)
  {
    // the other buffer set: its geometry half may start as soon as the raster half that last read the set (two frames ago)
    // is done -- normally long ago -- i.e. while the previous frame's raster half is still running
    const int set = c->par ^ 1;
    if(c->rasterRecorded[set])
      CUDA_TRY(c, cudaStreamWaitEvent(c->geoStream, c->evRasterDone[set], 0));
    if(c->mainDirty)
    {
      // a scene upload or stage-by-stage work is queued on the main stream: the geometry stream gets in line behind it
      CUDA_TRY(c, cudaEventRecord(c->evMain, c->stream));
      CUDA_TRY(c, cudaStreamWaitEvent(c->geoStream, c->evMain, 0));
      c->mainDirty = false;
    }
    selectSet(c, set);
    if((r = uploadSceneData(c, ubo, c->geoStream)) != OIT_OK)
      return r;
  }
  else if((r = uploadSceneData(c, ubo, c->stream)) != OIT_OK)
    return r;
  if((r = enqueueFrame(c)) != OIT_OK)
    return r;
  c->framePending = true;
  // asynchronous by default: the frame (and the frames before it) may still be running when this returns; oit_synchronize,
  // oit_download / oit_read_color, oit_get_stats and every scene change complete it first.  The NCCL gather's first,
  // un-captured frame and OIT_B200_SYNC_RENDER=1 wait here.
  if(!c->asyncRender || (c->gather && !c->gatherWarm))
    return finishFrame(c);
  return OIT_OK;
}

int oit_get_stats(OitCtx* c, OitStats* out)
{
  if(!c || !out)
    return OIT_ERR_INVALID_ARG;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  unsigned long long h[NUM_STAT_SLOTS];
  CUDA_TRY(c, cudaMemcpy(h, c->stats[c->par].p, sizeof(h), cudaMemcpyDeviceToHost));
  OitStats s{};
  s.fragments         = h[STAT_FRAGMENTS];
  s.fragmentsStored   = h[STAT_STORED];
  s.fragmentsTail     = h[STAT_TAIL];
  s.opaqueFragments   = h[STAT_OPAQUE];
  s.trianglesRejected = h[STAT_REJECTED];
  uint32_t nt, no;
  splitObjects(c, nt, no);
  s.trianglesDrawn = (uint64_t)nt * (c->idxPerObj / 3);
  s.tilePairs      = (uint64_t)c->pairTotal[0] + c->pairTotal[1];
  s.kernelLaunches = c->launches;
  if(c->counter.p)
  {
    uint32_t v = 0;
    CUDA_TRY(c, cudaMemcpy(&v, c->counter.p, 4, cudaMemcpyDeviceToHost));
    s.llCounter = v;
  }
  auto ms = [&](int a, int b) {
    float t = 0.f;
    if(c->evRecorded[a] && c->evRecorded[b] && cudaEventElapsedTime(&t, c->ev[a], c->ev[b]) == cudaSuccess)
      return t;
    cudaGetLastError();  // do not leave a sticky error behind for the host application
    return 0.f;
  };
  s.msGeometry  = ms(EV_START, EV_GEOM);
  s.msClear     = ms(EV_RASTER_START, EV_CLEAR);
  s.msOpaque    = ms(EV_CLEAR, EV_OPAQUE);
  s.msColor     = ms(EV_OPAQUE, EV_COLOR);
  s.msComposite = ms(EV_COLOR, EV_COMPOSITE);
  s.msResolve   = ms(EV_COMPOSITE, EV_RESOLVE);
  // the two halves of a frame run on two streams (the next frame's geometry overlaps this frame's raster): their sum
  s.msFrame     = s.msGeometry + ms(EV_RASTER_START, c->evRecorded[EV_RESOLVE] ? EV_RESOLVE : EV_COLOR);
  s.msExchangeWait = (float)((double)h[STAT_WAIT_NS] * 1e-6);
  c->lastStats  = s;
  *out          = s;
  return OIT_OK;
}

int oit_buffer_size(const OitCtx* c, OitBuffer which, size_t* bytes)
{
  if(!c || !bytes)
    return OIT_ERR_INVALID_ARG;
  DevBuf* b = bufferOf(const_cast<OitCtx*>(c), which);
  if(!b)
    return OIT_ERR_INVALID_ARG;
  *bytes = b->p ? b->bytes : 0;
  return OIT_OK;
}

int oit_download(OitCtx* c, OitBuffer which, void* host, size_t bytes)
{
  if(!c || !host)
    return OIT_ERR_INVALID_ARG;
  DevBuf* b = bufferOf(c, which);
  if(!b || !b->p)
    return fail(c, OIT_ERR_INVALID_ARG, "buffer not allocated for this configuration");
  if(bytes != b->bytes)
    return fail(c, OIT_ERR_SIZE, "host size does not match the device buffer");
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  CUDA_TRY(c, cudaMemcpyAsync(host, b->p, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return OIT_OK;
}

int oit_upload(OitCtx* c, OitBuffer which, const void* host, size_t bytes)
{
  if(!c || !host)
    return OIT_ERR_INVALID_ARG;
  DevBuf* b = bufferOf(c, which);
  if(!b || !b->p)
    return fail(c, OIT_ERR_INVALID_ARG, "buffer not allocated for this configuration");
  if(bytes != b->bytes)
    return fail(c, OIT_ERR_SIZE, "host size does not match the device buffer");
  c->mainDirty = true;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  CUDA_TRY(c, cudaMemcpyAsync(b->p, host, bytes, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return OIT_OK;
}

void* oit_device_ptr(OitCtx* c, OitBuffer which)
{
  if(!c)
    return nullptr;
  DevBuf* b = bufferOf(c, which);
  return b ? b->p : nullptr;
}

int oit_read_color(OitCtx* c, void* bgra8, size_t bytes) { return oit_download(c, OIT_BUF_FINAL, bgra8, bytes); }

void* oit_stream(OitCtx* c) { return c ? (void*)c->stream : nullptr; }

int oit_band_gather_unique_id(void* id128)
{
  if(!id128)
    return OIT_ERR_INVALID_ARG;
  std::string err;
  const int   r = gatherUniqueId(id128, err);
  if(r != OIT_OK)
    g_createError = err;
  return r;
}

int oit_enable_band_gather(OitCtx* c, const void* id128)
{
  if(!c || !id128)
    return OIT_ERR_INVALID_ARG;
  if(c->gather)
    return OIT_OK;
  if(c->peers)
    return fail(c, OIT_ERR_INVALID_ARG, "the peer-memory exchange is already enabled on this context");
  if(c->cfg.width % 4)
    return fail(c, OIT_ERR_INVALID_ARG, "the band gather needs a width that is a multiple of 4");
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  // rows of the largest band: every rank contributes a slice of that many (padded) rows
  uint32_t pad = 0;
  for(uint32_t b = 0; b < c->cfg.bandCount; b++)
  {
    uint32_t rows = 0;
    for(uint32_t y0 = 0; y0 < c->cfg.height; y0 += c->stripRows)
      if((y0 / c->stripRows) % c->cfg.bandCount == b)
        rows += std::min(c->stripRows, c->cfg.height - y0);
    pad = std::max(pad, rows);
  }
  c->padRows = pad;
  const size_t slice = (size_t)(pad + 1) * c->cfg.width;  // + one metadata row (the band's overflow flag, see oit_gather.cu)
  int          r;
  if((r = devAlloc(c, c->gatherBuf, slice * c->cfg.bandCount * 4)) != OIT_OK)
    return r;
  if((r = devAlloc(c, c->frame, (size_t)c->cfg.width * c->cfg.height * 4)) != OIT_OK)
    return r;
  CUDA_TRY(c, cudaMemset(c->gatherBuf.p, 0, c->gatherBuf.bytes));
  c->gather = gatherCreate(id128, (int)c->cfg.bandIndex, (int)c->cfg.bandCount, c->error);
  if(!c->gather)
    return OIT_ERR_CUDA;
  // the resolve now writes straight into this rank's slice of the gather buffer (in-place all-gather)
  if(c->finOwned)
    devFree(c->fin);
  c->finOwned   = false;
  c->fin.p      = (uint32_t*)c->gatherBuf.p + slice * c->cfg.bandIndex;
  c->fin.bytes  = std::max<size_t>((size_t)c->cfg.width * c->localOutH, 1) * 4;
  c->fp.fin     = (uint32_t*)c->fin.p;
  c->graphValid = false;
  c->gatherWarm = false;
  return OIT_OK;
}

int oit_band_peer_export(OitCtx* c, void* handle64)
{
  if(!c || !handle64)
    return OIT_ERR_INVALID_ARG;
  if(c->gather)
    return fail(c, OIT_ERR_INVALID_ARG, "the NCCL band gather is already enabled on this context");
  if(c->cfg.width % 4)
    return fail(c, OIT_ERR_INVALID_ARG, "the split-frame exchange needs a width that is a multiple of 4");
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  if(c->peers)
    return fail(c, OIT_ERR_INVALID_ARG, "oit_band_peer_export has already been called");
  const size_t bytes = (size_t)c->cfg.width * c->cfg.height * 4;
  c->peers           = peerCreate((int)c->cfg.bandIndex, (int)c->cfg.bandCount, bytes, handle64, c->error);
  c->peersOpen       = false;
  c->peersUnmapped   = false;
  c->peerSeq         = 0;  // (the new buffers' device-side frame counter starts at 0 too)
  if(!c->peers)
    return OIT_ERR_UNSUPPORTED;
  devFree(c->frame);
  c->frame.p     = peerFrame(c->peers, c->peerSeq);
  c->frame.bytes = bytes;
  return OIT_OK;
}

int oit_band_peer_enable(OitCtx* c, const void* handles, uint32_t count)
{
  if(!c || !handles)
    return OIT_ERR_INVALID_ARG;
  if(!c->peers || count != c->cfg.bandCount)
    return fail(c, OIT_ERR_INVALID_ARG, "oit_band_peer_enable: call oit_band_peer_export first and pass one handle per band");
  if(c->peersOpen)
    return OIT_OK;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  const int r = peerOpen(c->peers, handles, c->error);
  if(r != OIT_OK)
    return r;
  c->peersOpen     = true;
  c->peersUnmapped = false;
  c->graphValid    = false;
  // the queue of finished tiles for the pusher CTAs (+ its tail counter); entries are consumed and zeroed every frame
  c->pushers = c->cfg.bandCount >= 4 ? 16 : 0;
  if(const char* e = getenv("OIT_B200_PUSHERS"))
    c->pushers = std::max(0, std::min(64, atoi(e)));
  const size_t words = (size_t)c->fp.tilesX * c->fp.tileRowsLocal + 1;
  if(c->pushQueue.bytes != words * 4)
  {
    const int rq = devAlloc(c, c->pushQueue, words * 4);
    if(rq != OIT_OK)
      return rq;
  }
  CUDA_TRY(c, cudaMemset(c->pushQueue.p, 0, words * 4));
  return OIT_OK;
}

int oit_band_peer_disable(OitCtx* c)
{
  if(!c)
    return OIT_ERR_INVALID_ARG;
  if(!c->peers)
    return OIT_OK;
  CUDA_TRY(c, cudaSetDevice(c->cfg.device));
  if(const int fr = finishFrame(c); fr != OIT_OK)  // completes a frame oit_render left in flight, then the stream is idle
    return fr;
  c->graphValid = false;
  if(!c->peersUnmapped)
  {
    // first call, on every band and followed by a host barrier: the other bands' buffers are unmapped (if this band got as
    // far as mapping them); nothing is freed yet, because other bands may still have THIS band's buffer mapped
    peerClose(c->peers);
    c->peersOpen     = false;
    c->peersUnmapped = true;
    return OIT_OK;
  }
  peerDestroy(c->peers);  // second call: the exported buffer itself goes
  c->peers         = nullptr;
  c->peersOpen     = false;
  c->peersUnmapped = false;
  c->frame         = DevBuf{};
  return OIT_OK;
}

int oit_local_row_to_global(const OitCtx* c, uint32_t localRow, uint32_t* globalRow)
{
  if(!c || !globalRow || localRow >= c->localOutH)
    return OIT_ERR_INVALID_ARG;
  const uint32_t strip = localRow / c->stripRows, within = localRow % c->stripRows;
  *globalRow           = (strip * c->cfg.bandCount + c->cfg.bandIndex) * c->stripRows + within;
  return OIT_OK;
}

}  // extern "C"
