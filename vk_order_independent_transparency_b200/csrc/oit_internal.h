// oit_internal.h -- structures shared by the host API and the CUDA translation units of liboit_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <string>

#include "../../include/oit_b200.h"

namespace oit {

// Screen tiles: one CTA owns one tile for a whole geometry pass, so every per-pixel structure (A-buffer slice,
// aux words, colour and depth samples) is touched by exactly one SM and primitive order per pixel can be kept.
constexpr int TILE_W      = 16;
constexpr int TILE_H      = 16;
constexpr int TILE_PIX    = TILE_W * TILE_H;
constexpr int TILE_SHIFT  = 4;
#ifndef OIT_RASTER_THREADS
#define OIT_RASTER_THREADS 256
#endif
constexpr int RASTER_THREADS = OIT_RASTER_THREADS;  // = triangles staged per chunk
constexpr int SUBPIXEL_BITS  = 8;    // fixed-point snapping (NVIDIA and lavapipe use 8)
constexpr float GUARD_BAND_PX = 2097152.0f;  // |x|,|y| < 2^21 px so every edge delta fits in 31 bits

// raster kernel variants
enum RasterPass
{
  PASS_SIMPLE = 0,
  PASS_LINKEDLIST,
  PASS_LOOP_COLOR,
  PASS_LOOP64,
  PASS_SPINLOCK,
  PASS_INTERLOCK,
  PASS_WEIGHTED,
  PASS_LOOP_DEPTH,
  PASS_OPAQUE,
  NUM_RASTER_PASSES
};

// post-vertex-stage vertex: object.vert.glsl:32-38 + viewport transform, snapped to 1/256 px
struct __align__(16) TVert
{
  int32_t x, y;   // x == INT32_MIN marks a vertex outside the clip volume / guard band
  float   z;      // z_ndc = z_clip / w_clip
  float   invw;
};
static_assert(sizeof(TVert) == 16, "TVert layout: one 128-bit load per vertex");
// What the shading interpolates (shaderCommon.glsl:25-31), repacked by the vertex stage into two 128-bit words per vertex
// behind the TVert table (FrameParams::tvAttr): [2i] = normal.xyz, colour.r; [2i+1] = colour.gba, (viewMatrix * pos).z
// (Interpolants.depth, only WBOIT reads it).  One 32-byte sector per vertex instead of 28 bytes that straddle two sectors of
// the 40-byte vertex record, and 2 loads instead of 4.
constexpr int ATTR_FLOATS = 8;

// One piece of a near-clipped triangle (oit_clip.cuh), written by the binning of the frame and read by the raster kernel:
// the piece's post-projection vertices and, per vertex, where its attributes come from.
struct ClipEntry
{
  TVert v[3];                  // post-projection vertices of the piece
  float attr[3][ATTR_FLOATS];  // their attribute records in the layout of FrameParams::tvAttr (normal3, colour4, view depth):
                               // an original vertex's, or the clip-space interpolation fma(t, attr[Q] - attr[P], attr[P])
                               // along the cut edge -- so that the shading reads either through the same code
};
static_assert(offsetof(ClipEntry, attr) == 48 && sizeof(ClipEntry) == 144, "ClipEntry layout: 128-bit attribute loads");
constexpr uint32_t PAIR_CLIPPED = 0x80000000u;  // pair value: bit 31 set = index of a ClipEntry instead of a triangle
constexpr uint32_t PAIR_SKIP    = 0xFFFFFFFFu;  // a piece that found no room in the entry table (the frame is rendered again)

enum StatSlot
{
  STAT_FRAGMENTS = 0,
  STAT_STORED,
  STAT_TAIL,
  STAT_OPAQUE,
  STAT_REJECTED,
  STAT_OVERFLOW,  // the (tile, triangle) pair buffer of a draw was too small: the frame must be rendered again
  STAT_PEER_TIMEOUT,  // split frame over peer memory: a band did not arrive at the frame barrier in time
  STAT_SCRATCH,       // scene upload: index validation flag (no frame is in flight then)
  STAT_INTERNAL,      // a look-back chain of the binning gave up (cannot happen; reported instead of hanging the GPU)
  STAT_OVERFLOW_ANY,  // split frame: SOME band raised STAT_OVERFLOW for this frame (carried by the exchange itself), so that
                      // every band takes the same decision to render the frame again
  STAT_WAIT_NS,       // split frame: nanoseconds this band spent in the READY and DONE waits of the frame
  NUM_STAT_SLOTS = 12
};

// Split frame over NVLink peer memory (oit_peer.cu): band b's resolved pixels are stored straight into the whole-frame
// buffer of every band (its own included) by the kernel that produces them.
constexpr int PEER_MAX       = 16;
// Every band has TWO whole-frame buffers; frame n of the context goes to buffer n & 1 of every band, so that the bands only
// have to agree on frame boundaries one frame late (see oit_peer.cu).
constexpr int PEER_FLAG_READY = 0;             // flags[READY + b] = n: band b allows frame n to be written into ITS buffer n & 1
constexpr int PEER_FLAG_DONE  = PEER_MAX;      // flags[DONE + b]  = n: band b's strips of frame n are in THIS band's buffer n & 1
constexpr int PEER_FLAG_SEQ   = 2 * PEER_MAX;  // frames this band has issued completely (local use); the running frame is SEQ + 1
constexpr int PEER_FLAG_OVF   = 2 * PEER_MAX + 1;  // flags[OVF + (n & 1) * PEER_MAX + b] = n: band b's frame n overflowed a pair / clip buffer
constexpr int PEER_FLAG_WORDS = 96;
static_assert(PEER_FLAG_OVF + 2 * PEER_MAX <= PEER_FLAG_WORDS, "peer flag page");
struct PeerTable
{
  uint32_t* frame[2][PEER_MAX];  // whole frames [H][W] BGRA8 of band b (peer-mapped, own entry local)
  uint32_t* flags[PEER_MAX];     // PEER_FLAG_WORDS words behind them
};
// the frame buffer of every band the running frame is written to (device side)
__device__ __forceinline__ uint32_t* const* peerFramesOfRunningFrame(const PeerTable* t, int bandIndex)
{
  const uint32_t n = *reinterpret_cast<volatile const uint32_t*>(t->flags[bandIndex] + PEER_FLAG_SEQ) + 1u;
  return t->frame[n & 1u];
}

struct DeviceUbo
{
  float projView[16];
  float view[16];
  float alphaMin, alphaWidth;
  float pad[2];
};

// Bucket table of the sRGB encoder: the linear values in [2^-13, 1] are cut into buckets by the top 16 bits of their float
// representation (128 buckets per octave); a bucket is narrower than the distance between two encode thresholds, so it
// holds the code of its lower end and at most one threshold (checked when the table is built, oit_api.cu: buildTables).
constexpr uint32_t SRGB_BUCKET_BASE  = 114u << 7;   // bits(2^-13) >> 16; everything below encodes to 0 (thr[1] = 1.5e-4)
constexpr uint32_t SRGB_BUCKET_COUNT = 1665u;       // up to and including bits(1.0f) >> 16
constexpr uint32_t SRGB_BUCKET_BYTES = 1680u;       // padded to a multiple of 16
constexpr uint32_t SRGB_TABLE_FLOATS = 256u + 260u + 256u;
constexpr uint32_t SRGB_TABLE_BYTES  = SRGB_TABLE_FLOATS * 4u + SRGB_BUCKET_BYTES;


struct FrameParams
{
  // render target (after supersample); W x H is the full frame, localH the rows this band owns
  int      W, H, localH;
  int      msaa, sampleShading, coverage;
  int      L;
  uint32_t capacity;  // linked list pool size in nodes (scene.linkedListAllocatedPerElement)
  int      tailBlend;
  int      layers;    // A-buffer / aux layers (msaa if sample shading)
  uint32_t clearColor;  // (0.2,0.2,0.2,0.2) linear encoded to BGRA8 sRGB (oitRender.cpp:90)
  int      algorithm;   // OIT_*
  int      supersample; // 1 or 2
  int      fused;       // the colour-pass kernel also composites and resolves its tile (oit_render fast path)
  int      onChip;      // fused only: the tile's A-buffer slice + aux words live in shared memory (k-buffer techniques)
  // tiles
  int tilesX, tileRowsGlobal, tileRowsLocal;
  int stripTileRows, bandCount, bandIndex;
  const int32_t* rowLocal;  // [tileRowsGlobal] local tile row of a global tile row this band owns, -1 for another band's row
  // the per-frame part of shaderio::SceneData, in device memory so that a captured frame graph can be replayed
  const DeviceUbo* ubo;
  // buffers (device)
  uint32_t*            abuf;
  uint32_t*            aux;
  uint32_t*            spin;
  uint32_t*            adepth;
  uint32_t*            counter;
  uint32_t*            color;
  float*               depth;  // nullptr when nothing opaque is drawn: every sample reads 1.0
  uint16_t*            wacc;
  uint16_t*            wrev;
  uint32_t*            fin;
  const PeerTable*     peers;   // split frame over peer memory: every band's whole-frame buffer (nullptr = off)
  // ... with pusher CTAs (oit_raster_ll.cu): the tile CTAs only write `fin` and publish the finished tile in pushQueue; the
  // first `pushers` CTAs of the grid copy finished tiles into every band's frame with 128-bit stores, so that NVLink
  // back-pressure stalls THEM and not the SMs that rasterise.  0 = the tile's own threads store to the peers.
  uint32_t*            pushQueue;  // [numLocalTiles] tile + 1 in completion order (0 = not yet), [numLocalTiles] = the tail counter
  int                  pushers;
  const float*         tables;  // the SrgbTables image (oit_device.cuh), SRGB_TABLE_BYTES
  unsigned long long*  stats;
  // geometry
  const float*    verts;
  const uint32_t* indices;
  TVert*          tv;
  float4*         tvAttr;   // [2 * nVerts] behind the TVert table: the vertices' attribute records (see ATTR_FLOATS)
  uint32_t        nVerts;
  // binning of the current draw
  const uint32_t* pairTri;    // triangle index (first index / 3) per (tile, triangle) pair, tile-major, in order
  const uint32_t* tileStart;  // [numLocalTiles + 1]
  const uint32_t* tileOrder;  // [numLocalTiles] tile handled by CTA i: heaviest triangle lists first
  ClipEntry*      clipEntries;  // pieces of the near-clipped triangles of the current draw
  uint32_t        clipCapacity;
};

// Fused frame kernel, k-buffer techniques without sample shading: the tile's A-buffer slice and aux words can live in
// shared memory for the whole pass ("on chip").  Words needed per tile: [A-buffer][imgAux][imgDepth][imgSpin]; 0 = the
// technique keeps its A-buffer in HBM (linked list: unbounded; Loop32: two geometry passes; WBOIT: no A-buffer).
__host__ __device__ inline uint32_t onChipAbufWords(int algorithm, int L, int coverage)
{
  if(algorithm == OIT_LOOP64)
    return (uint32_t)TILE_PIX * L * 2u;
  if(algorithm == OIT_SIMPLE || algorithm == OIT_SPINLOCK || algorithm == OIT_INTERLOCK)
    return (uint32_t)TILE_PIX * L * (coverage ? 4u : 2u);
  return 0u;
}
__host__ __device__ inline uint32_t onChipWords(int algorithm, int L, int coverage)
{
  const uint32_t a = onChipAbufWords(algorithm, L, coverage);
  return a ? a + 3u * TILE_PIX : 0u;
}
constexpr uint32_t ON_CHIP_MAX_BYTES = 36u * 1024u;

__host__ __device__ inline int tileRowOwner(int R, int stripTileRows, int bandCount) { return (R / stripTileRows) % bandCount; }
__host__ __device__ inline int tileRowToLocal(int R, int stripTileRows, int bandCount)
{
  return (R / (stripTileRows * bandCount)) * stripTileRows + (R % stripTileRows);
}
__host__ __device__ inline int tileRowToGlobal(int r, int stripTileRows, int bandCount, int bandIndex)
{
  return ((r / stripTileRows) * bandCount + bandIndex) * stripTileRows + (r % stripTileRows);
}

// ---- launchers (each returns the number of kernels it launched) -------------------------------------------------
int launchTransformVertices(const FrameParams& p, cudaStream_t s);
void launchExpandSpheres(const float* spheres, uint32_t nSpheres, const float* unitPos, uint32_t vPer, const uint32_t* unitTri, uint32_t iPer,
                         float* verts, uint32_t* indices, cudaStream_t s);
void launchValidateIndices(const uint32_t* indices, uint32_t n, uint32_t nVerts, unsigned long long* bad, cudaStream_t s);
// bins triangles [firstTri, firstTri+triCount) of the index buffer; cullBack for the opaque draw.
// d_counts/d_offsets: triCount+1 words; returns pair total through *hTotal (synchronises the stream once).
struct BinBuffers
{
  uint32_t* lb;          // look-back state of the binning: [pairInfo 4 words][tickets 4 words][descriptors]; zeroed every frame
  size_t    lbBytes;
  uint32_t* pairKey[2];  // [pairCapacity]
  uint32_t* pairVal[2];
  uint32_t* tileStart;   // [numLocalTiles + 1]
  uint32_t* pairInfo;    // = lb: [0] pairs present, [1] pairs wanted, [2] clip entries wanted (read back with the statistics)
  uint32_t* tileOrder;   // [numLocalTiles] launch order of the tiles: heaviest lists first
  ClipEntry* clipEntries;  // [clipCapacity]; pairInfo[2] = entries wanted by the last frame
  size_t     clipCapacity;
  uint32_t* scratch;     // scan / histogram scratch
  size_t    scratchWords;
  size_t    pairCapacity;
  size_t    triCapacity;
};
int launchBin(const FrameParams& p, const BinBuffers& b, uint32_t firstTri, uint32_t triCount, bool cullBack, int* sortedBuf,
              cudaStream_t s);
size_t binScratchWords(size_t triCount, size_t pairCapacity, size_t numTiles);
size_t binLookbackBytes(size_t triCount, size_t pairCapacity);

// band gather of the split-frame mode (oit_gather.cu); NCCL is loaded lazily with dlopen
struct BandGatherState;
int              gatherUniqueId(void* id128, std::string& err);
BandGatherState* gatherCreate(const void* id128, int rank, int world, std::string& err);
void             gatherDestroy(BandGatherState* g);
int gatherLaunch(BandGatherState* g, uint32_t* gathered, uint32_t* frame, int W, int H, int stripRows, int padRows, unsigned long long* stats,
                 cudaStream_t s, std::string& err);

// split frame over peer memory (oit_peer.cu): CUDA IPC mappings of every band's frame buffer + flag barrier kernels
struct PeerState;
PeerState*       peerCreate(int rank, int world, size_t frameBytes, void* handle64, std::string& err);
int              peerOpen(PeerState* ps, const void* handles, std::string& err);  // world x 64 bytes, in band order
void             peerClose(PeerState* ps);                                         // unmaps the other bands' buffers
void             peerDestroy(PeerState* ps);
uint32_t*        peerFrame(PeerState* ps, unsigned which);  // this band's frame buffer `which` (0 / 1)
const PeerTable* peerTable(PeerState* ps);
// first / last node of a frame's raster half, and the wait that completes the latest frame for the host (oit_peer.cu)
int peerFrameBegin(PeerState* ps, unsigned long long* stats, uint32_t* zeroA, uint32_t* zeroB, cudaStream_t s);
int peerFrameEnd(PeerState* ps, unsigned long long* stats, unsigned long long* hostMirror, int mirrorWords, const uint32_t* pairInfoA,
                 const uint32_t* pairInfoB, cudaStream_t s);
int peerFrameFlush(PeerState* ps, unsigned long long* stats, unsigned long long* hostMirror, cudaStream_t s);
int peerScatterRows(PeerState* ps, const uint32_t* fin, int W, int localRows, int stripRows, cudaStream_t s);

int launchClears(const FrameParams& p, int algorithm, cudaStream_t s, bool skipListHeads = false, bool skipCounter = false);
int launchRaster(const FrameParams& p, int pass, cudaStream_t s);
int launchRasterLinkedList(const FrameParams& p, cudaStream_t s);  // oit_raster_ll.cu
bool linkedListFrameStartsEmpty(const FrameParams& p);             // oit_raster.cu: the fused frame goes to k_raster_ll
int launchRasterQueued(const FrameParams& p, int pass, cudaStream_t s);  // oit_raster_q.cu
int launchComposite(const FrameParams& p, int algorithm, cudaStream_t s);
int launchResolve(const FrameParams& p, int supersample, int outW, int outLocalH, cudaStream_t s);

}  // namespace oit
