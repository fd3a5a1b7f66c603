// oit_geometry.cu -- vertex stage, triangle setup / tile binning, and the small scan + stable radix sort they need.
//
// Replaces: object.vert.glsl:32-38 (K0), the fixed-function viewport transform and primitive assembly
// (main.cpp:504-532) of the reference.  Output: per (local) screen tile, the list of triangles whose sample
// bounding box touches it, IN PRIMITIVE ORDER -- the order the ROP and the ordered interlock rely on.
//
// Everything here is asynchronous: the number of (tile, triangle) pairs stays on the device (the sort and the range
// kernels read it from memory, grids are sized for the buffer capacity), so a whole frame can be replayed as one CUDA
// graph.  If the pair buffer is too small an overflow flag is raised and the host grows it and renders again.
#include "oit_clip.cuh"
#include "oit_device.cuh"

namespace oit {

// ------------------------------------------------------------------------------------------------------------------
// vertex stage
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_transform_vertices(const FrameParams p)
{
  const float  hw = 0.5f * (float)p.W, hh = 0.5f * (float)p.H;
  const float* M = p.ubo->projView;
  const float* V = p.ubo->view;
  for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < p.nVerts; i += gridDim.x * blockDim.x)
  {
    const float* v  = p.verts + (size_t)i * 10;
    const float  px = v[0], py = v[1], pz = v[2];
    float        clip[4];
#pragma unroll
    for(int r = 0; r < 4; r++)
      clip[r] = __fmaf_rn(M[0 + r], px, __fmaf_rn(M[4 + r], py, __fmaf_rn(M[8 + r], pz, M[12 + r])));
    const float viewz = __fmaf_rn(V[2], px, __fmaf_rn(V[6], py, __fmaf_rn(V[10], pz, V[14])));
    p.tv[i]      = finishVertex(clip, hw, hh);
    // the interpolants of the shading, repacked (oit_internal.h: ATTR_FLOATS)
    p.tvAttr[2 * (size_t)i]     = make_float4(v[3], v[4], v[5], v[6]);
    p.tvAttr[2 * (size_t)i + 1] = make_float4(v[7], v[8], v[9], viewz);
  }
}

// scene upload: *bad |= 1 if any index refers to a vertex that does not exist (checked where the data already is)
__global__ void __launch_bounds__(256) k_validate_indices(const uint32_t* __restrict__ indices, uint32_t n, uint32_t nVerts, unsigned long long* bad)
{
  bool any = false;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    any |= indices[i] >= nVerts;
  if(__syncthreads_or(any) && threadIdx.x == 0)
    atomicOr(bad, 1ull);
}
void launchValidateIndices(const uint32_t* indices, uint32_t n, uint32_t nVerts, unsigned long long* bad, cudaStream_t s)
{
  const int grid = (int)min((size_t)148 * 8, ((size_t)n + 255) / 256);
  k_validate_indices<<<grid, 256, 0, s>>>(indices, n, nVerts, bad);
}

// instanced scene input: flattens numObjects copies of the unit sphere on the device, with the arithmetic of
// the host generator (oit_scene.cpp: pos = unit * radius + centre as a separate multiply and add), 40-byte vertices
__global__ void __launch_bounds__(256) k_expand_spheres(const float* __restrict__ spheres, uint32_t nSpheres, const float* __restrict__ unitPos,
                                                        uint32_t vPer, const uint32_t* __restrict__ unitTri, uint32_t iPer,
                                                        float* __restrict__ verts, uint32_t* __restrict__ indices)
{
  const size_t nV = (size_t)nSpheres * vPer, nI = (size_t)nSpheres * iPer;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nV + nI; i += (size_t)gridDim.x * blockDim.x)
  {
    if(i < nV)
    {
      const uint32_t obj = (uint32_t)(i / vPer), v = (uint32_t)(i - (size_t)obj * vPer);
      const float4   a = __ldg(reinterpret_cast<const float4*>(spheres) + 2 * (size_t)obj);      // centre, radius
      const float4   c = __ldg(reinterpret_cast<const float4*>(spheres) + 2 * (size_t)obj + 1);  // colour
      const float    px = unitPos[3 * v], py = unitPos[3 * v + 1], pz = unitPos[3 * v + 2];
      float2*        d  = reinterpret_cast<float2*>(verts + i * 10);  // 40-byte stride: 8-byte aligned
      d[0]              = make_float2(__fadd_rn(__fmul_rn(px, a.w), a.x), __fadd_rn(__fmul_rn(py, a.w), a.y));
      d[1]              = make_float2(__fadd_rn(__fmul_rn(pz, a.w), a.z), px);
      d[2]              = make_float2(py, pz);
      d[3]              = make_float2(c.x, c.y);
      d[4]              = make_float2(c.z, c.w);
    }
    else
    {
      const size_t   j   = i - nV;
      const uint32_t obj = (uint32_t)(j / iPer), k = (uint32_t)(j - (size_t)obj * iPer);
      indices[j]         = obj * vPer + unitTri[k];
    }
  }
}
void launchExpandSpheres(const float* spheres, uint32_t nSpheres, const float* unitPos, uint32_t vPer, const uint32_t* unitTri, uint32_t iPer,
                         float* verts, uint32_t* indices, cudaStream_t s)
{
  const size_t total = (size_t)nSpheres * vPer + (size_t)nSpheres * iPer;
  const int    grid  = (int)min((size_t)148 * 16, (total + 255) / 256);
  k_expand_spheres<<<grid, 256, 0, s>>>(spheres, nSpheres, unitPos, vPer, unitTri, iPer, verts, indices);
}

int launchTransformVertices(const FrameParams& p, cudaStream_t s)
{
  if(p.nVerts == 0)
    return 0;
  const int blocks = (int)min((p.nVerts + 255u) / 256u, 148u * 16u);
  k_transform_vertices<<<blocks, 256, 0, s>>>(p);
  return 1;
}

// ------------------------------------------------------------------------------------------------------------------
// triangle -> tile range
// ------------------------------------------------------------------------------------------------------------------
struct TileRange
{
  int tx0, tx1, ty0, ty1;  // global tile coordinates, inclusive
};

// The pixel range that can contain a covered sample: a sample of pixel px lies at px*256 + off, off in [lo, hi].
// what the tile-range / band logic reads, BY VALUE for the out-of-line clipped path (see ClipInput)
struct BinView
{
  ClipInput      in;
  int            msaa, tilesX;
  const int32_t* rowLocal;  // [tileRowsGlobal] local tile row of a global tile row this band owns, -1 for the others' rows;
                            // nullptr when a single band owns every row (no look-ups, no integer divisions per triangle row)
};
__device__ __forceinline__ BinView binView(const FrameParams& p)
{
  return BinView{clipInput(p), p.msaa, p.tilesX, p.bandCount > 1 ? p.rowLocal : nullptr};
}
__device__ __forceinline__ bool tvTileRange(const BinView& v, const TVert& a, const TVert& b, const TVert& c, bool cullBack, TileRange& r)
{
  const long long area2 = (long long)(b.x - a.x) * (c.y - a.y) - (long long)(c.x - a.x) * (b.y - a.y);
  if(area2 == 0 || (cullBack && area2 > 0))
    return false;
  const int lo = v.msaa == 1 ? 128 : (v.msaa == 4 ? 32 : 16), hi = 256 - lo;
  const int minx = min(a.x, min(b.x, c.x)), maxx = max(a.x, max(b.x, c.x));
  const int miny = min(a.y, min(b.y, c.y)), maxy = max(a.y, max(b.y, c.y));
  const int px0 = max((minx - hi + 255) >> 8, 0), px1 = min((maxx - lo) >> 8, v.in.W - 1);
  const int py0 = max((miny - hi + 255) >> 8, 0), py1 = min((maxy - lo) >> 8, v.in.H - 1);
  if(px0 > px1 || py0 > py1)
    return false;
  r.tx0 = px0 >> TILE_SHIFT;
  r.tx1 = px1 >> TILE_SHIFT;
  r.ty0 = py0 >> TILE_SHIFT;
  r.ty1 = py1 >> TILE_SHIFT;
  return true;
}
__device__ __forceinline__ uint32_t ownedTiles(const BinView& v, const TileRange& r)
{
  const int nx = r.tx1 - r.tx0 + 1;
  if(v.rowLocal == nullptr)
    return (uint32_t)(nx * (r.ty1 - r.ty0 + 1));
  uint32_t n = 0;
  for(int R = r.ty0; R <= r.ty1; R++)
    if(__ldg(v.rowLocal + R) >= 0)
      n += nx;
  return n;
}
__device__ __forceinline__ uint32_t emitTiles(const BinView& v, const TileRange& r, uint32_t val, uint32_t o, uint32_t* __restrict__ keys,
                                              uint32_t* __restrict__ vals)
{
  for(int R = r.ty0; R <= r.ty1; R++)
  {
    const int lr = v.rowLocal ? __ldg(v.rowLocal + R) : R;
    if(lr >= 0)
    {
      const uint32_t rowKey = (uint32_t)lr * v.tilesX;
      for(int tx = r.tx0; tx <= r.tx1; tx++, o++)
      {
        keys[o] = rowKey + tx;
        vals[o] = val;  // emitted in primitive order; the stable sort keeps it (and the pieces of a primitive) per tile
      }
    }
  }
  return o;
}

// A triangle with a vertex that has no post-projection position (oit_clip.cuh): the pairs of its pieces.  Out of line:
// rare, and the clipper's state stays out of the binning kernels' fast path.  Returns the pair count; *rejected = no piece.
static __device__ __noinline__ uint32_t countClipped(const BinView v, uint32_t i0, uint32_t i1, uint32_t i2, bool cullBack, bool* rejected)
{
  const ClipResult cr = clipTriangleNear(v.in, i0, i1, i2);
  *rejected           = cr.count == 0;
  uint32_t n          = 0;
  for(int s = 0; s < cr.count; s++)
  {
    TileRange r;
    if(tvTileRange(v, cr.v[s][0].v, cr.v[s][1].v, cr.v[s][2].v, cullBack, r))
      n += ownedTiles(v, r);
  }
  return n;
}
// ... and their emission: every piece that owns tiles gets a ClipEntry (vertices + vertex records), its pairs carry the
// entry's index instead of the triangle's
static __device__ __noinline__ void emitClipped(const BinView v, uint32_t i0, uint32_t i1, uint32_t i2, bool cullBack, uint32_t o,
                                                uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, ClipEntry* __restrict__ entries,
                                                uint32_t clipCapacity, uint32_t* __restrict__ clipCount, unsigned long long* __restrict__ overflow)
{
  const ClipResult cr    = clipTriangleNear(v.in, i0, i1, i2);
  const uint32_t   ix[3] = {i0, i1, i2};
  for(int s = 0; s < cr.count; s++)
  {
    TileRange r;
    if(!tvTileRange(v, cr.v[s][0].v, cr.v[s][1].v, cr.v[s][2].v, cullBack, r) || ownedTiles(v, r) == 0u)
      continue;
    const uint32_t e   = atomicAdd(clipCount, 1u);
    uint32_t       val = PAIR_SKIP;
    if(e < clipCapacity)
    {
      ClipEntry& ce = entries[e];
      for(int m = 0; m < 3; m++)
      {
        const ClipVert& cv = cr.v[s][m];
        ce.v[m]            = cv.v;
        const float* aP    = v.in.verts + (size_t)ix[cv.i] * 10 + 3;
        const float* aQ    = v.in.verts + (size_t)ix[cv.j] * 10 + 3;
        for(int c = 0; c < 7; c++)
          ce.attr[m][c] = cv.i == cv.j ? aP[c] : __fmaf_rn(cv.t, __fsub_rn(aQ[c], aP[c]), aP[c]);
        ce.attr[m][7] = cv.viewz;
      }
      val = PAIR_CLIPPED | e;
    }
    else
      atomicAdd(overflow, 1ull);  // the host grows the table and renders the frame again
    o = emitTiles(v, r, val, o, keys, vals);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// binning: two passes over the triangles, no chain between blocks
// ------------------------------------------------------------------------------------------------------------------
// Pass 1 leaves ONE number per block of BIN_BLOCK consecutive triangles (its pair count); in pass 2 every block sums the
// numbers of the blocks before it by itself (a few loads per thread), scans its own triangles and emits the (tile,
// triangle) pairs at their final offsets IN PRIMITIVE ORDER (the stable sort keeps it per tile).  Sizing a triangle is a
// handful of instructions on data that stays in L2, so doing it twice is cheaper than either storing per-triangle counts
// and scanning them, or chaining the blocks of a single pass through a look-back (measured: the blocks of a wave all wait
// for the wave's slowest loads, 45 us instead of 30).
constexpr int BIN_TPT   = 4;              // consecutive triangles per thread
constexpr int BIN_BLOCK = 256 * BIN_TPT;  // ... per block
constexpr int BIN_STAGE = 3072;           // pairs a block collects in shared memory before writing them out

// the tile range + pair count of the thread's j-th triangle (n = 0: nothing to emit)
struct BinItem
{
  TileRange r;
  uint32_t  n;
  bool      clipped;
};
__device__ __forceinline__ BinItem sizeTriangle(const FrameParams& p, const BinView& v, uint32_t tri, bool cullBack, uint32_t& nRejected)
{
  BinItem        it;
  const uint32_t i0 = p.indices[3 * (size_t)tri], i1 = p.indices[3 * (size_t)tri + 1], i2 = p.indices[3 * (size_t)tri + 2];
  const TVert    a = p.tv[i0], b = p.tv[i1], c = p.tv[i2];
  it.n       = 0;
  it.clipped = false;
  if(a.x != INT32_MIN && b.x != INT32_MIN && c.x != INT32_MIN)
  {
    if(tvTileRange(v, a, b, c, cullBack, it.r))
      it.n = ownedTiles(v, it.r);
  }
  else
  {
    bool rejected;
    it.n       = countClipped(v, i0, i1, i2, cullBack, &rejected);
    it.clipped = true;
    nRejected += rejected ? 1u : 0u;
  }
  return it;
}

__global__ void __launch_bounds__(256, 4) k_bin_count(const FrameParams p, uint32_t firstTri, uint32_t triCount, int cullBack,
                                                      uint32_t* __restrict__ blockTotals)
{
  __shared__ uint32_t sm[33];
  const BinView  v  = binView(p);
  const uint32_t t0 = blockIdx.x * BIN_BLOCK + threadIdx.x * BIN_TPT;
  uint32_t       sum = 0, nRejected = 0;
#pragma unroll
  for(int j = 0; j < BIN_TPT; j++)
    if(t0 + j < triCount)
      sum += sizeTriangle(p, v, firstTri + t0 + j, cullBack != 0, nRejected).n;
  uint32_t total;
  blockExclusiveScan(sum, sm, total);
  if(threadIdx.x == 0)
    blockTotals[blockIdx.x] = total;
  if(nRejected)
    atomicAdd(&p.stats[STAT_REJECTED], (unsigned long long)nRejected);
}

// Pairs beyond `capacity` are dropped and the overflow flag is raised (the host then grows the buffers and renders the
// frame again).  info: [0] pairs present, [1] pairs wanted, [2] clip entries handed out (zeroed before the binning)
__global__ void __launch_bounds__(256, 4) k_bin_emit(const FrameParams p, uint32_t firstTri, uint32_t triCount, int cullBack,
                                                     const uint32_t* __restrict__ blockTotals, uint32_t* __restrict__ keys,
                                                     uint32_t* __restrict__ vals, uint32_t capacity, uint32_t* __restrict__ info)
{
  __shared__ uint32_t sm[33];
  // pairs of all the blocks before this one
  uint32_t before = 0;
  for(uint32_t b = threadIdx.x; b < blockIdx.x; b += blockDim.x)
    before += blockTotals[b];
  const BinView  v  = binView(p);
  const uint32_t t0 = blockIdx.x * BIN_BLOCK + threadIdx.x * BIN_TPT;
  BinItem        it[BIN_TPT];
  uint32_t       sum = 0, unused = 0;
#pragma unroll
  for(int j = 0; j < BIN_TPT; j++)
  {
    it[j].n = 0;
    if(t0 + j < triCount)
      it[j] = sizeTriangle(p, v, firstTri + t0 + j, cullBack != 0, unused);
    sum += it[j].n;
  }
  uint32_t       base, blockTotal;
  blockExclusiveScan(before, sm, base);  // (its total is the sum over the threads: the pairs before this block)
  const uint32_t ex = blockExclusiveScan(sum, sm, blockTotal);
  // The block's pairs are collected in shared memory and leave as two coalesced streams (a thread's own pairs are a run of
  // one to four words: written directly, a warp's store touches up to 32 sectors); blocks with very large triangles write
  // directly.
  __shared__ uint32_t stageKeys[BIN_STAGE], stageVals[BIN_STAGE];
  const bool          staged = blockTotal <= (uint32_t)BIN_STAGE && base + blockTotal <= capacity;
  uint32_t*           dk     = staged ? stageKeys : keys;
  uint32_t*           dv     = staged ? stageVals : vals;
  uint32_t            o      = (staged ? 0u : base) + ex;
  const uint32_t      limit  = staged ? (uint32_t)BIN_STAGE : capacity;
#pragma unroll
  for(int j = 0; j < BIN_TPT; j++)
  {
    if(it[j].n && o + it[j].n <= limit)
    {
      const uint32_t tri = firstTri + t0 + j;
      if(!it[j].clipped)
        emitTiles(v, it[j].r, tri, o, dk, dv);
      else
        emitClipped(v, p.indices[3 * (size_t)tri], p.indices[3 * (size_t)tri + 1], p.indices[3 * (size_t)tri + 2], cullBack != 0, o, dk, dv,
                    p.clipEntries, p.clipCapacity, info + 2, p.stats + STAT_OVERFLOW);
    }
    o += it[j].n;
  }
  if(staged)
  {
    __syncthreads();
    for(uint32_t i = threadIdx.x; i < blockTotal; i += blockDim.x)
    {
      keys[base + i] = stageKeys[i];
      vals[base + i] = stageVals[i];
    }
  }
  if(blockIdx.x == gridDim.x - 1 && threadIdx.x == 0)
  {
    const uint32_t total = base + blockTotal;
    info[0]              = total > capacity ? 0u : total;  // pairs present (none when the frame has to be redone)
    info[1]              = total;                          // pairs wanted
    if(total > capacity)
      atomicAdd(&p.stats[STAT_OVERFLOW], 1ull);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// exclusive scan of uint32 in ONE pass: decoupled look-back (the histogram tables of the sort: a few dozen blocks)
// ------------------------------------------------------------------------------------------------------------------
// Blocks take a ticket with an atomic when they START, so every block with a smaller ticket is already running (or done)
// and never waits for a later one: the look-back cannot deadlock.  A descriptor is one 64-bit word -- state in the top two
// bits (0 = nothing yet, 1 = the block's own aggregate, 2 = inclusive prefix), value below -- written and read whole.
// The descriptors and tickets of a frame live in BinBuffers::lb, zeroed by one memset node at the start of the binning.
constexpr unsigned long long LB_AGGREGATE = 1ull << 62, LB_INCLUSIVE = 2ull << 62, LB_VALUE = (1ull << 62) - 1ull;
constexpr uint32_t           LB_SPIN_LIMIT = 1u << 22;  // ~0.1 s of polling: a bug must surface as an error, not as a hung GPU

__device__ __forceinline__ unsigned long long lbLoad(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void lbStore(unsigned long long* p, unsigned long long v)
{
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// The first WARP of the block with ticket `bid` (all 32 lanes): publishes the block's total and returns the sum of the totals
// of all blocks before it.  The lanes inspect 32 predecessors at a time.
__device__ __forceinline__ uint32_t lookbackExclusive(unsigned long long* desc, uint32_t bid, uint32_t total, unsigned long long* stats)
{
  const uint32_t lane = threadIdx.x & 31u;
  if(lane == 0)
    lbStore(desc + bid, (bid == 0 ? LB_INCLUSIVE : LB_AGGREGATE) | total);
  if(bid == 0)
    return 0u;
  uint32_t excl = 0, spins = 0;
  for(int window = (int)bid - 1;;)
  {
    const int                i     = window - (int)lane;  // lane 0 looks at the nearest predecessor
    const unsigned long long d     = i >= 0 ? lbLoad(desc + i) : LB_INCLUSIVE;  // (before block 0: an inclusive prefix of 0)
    const uint32_t           state = (uint32_t)(d >> 62);
    const uint32_t           inc = __ballot_sync(0xffffffffu, state == 2u), empty = __ballot_sync(0xffffffffu, state == 0u);
    const int                first = inc ? __ffs(inc) - 1 : 32;                       // nearest lane holding an inclusive prefix
    const uint32_t           need  = first >= 31 ? 0xffffffffu : ((2u << first) - 1u);  // lanes 0..first must have published
    if(empty & need)
    {
      if(++spins > LB_SPIN_LIMIT)
      {
        if(lane == 0)
          atomicAdd(stats + STAT_INTERNAL, 1ull);
        break;
      }
      __nanosleep(20);
      continue;
    }
    uint32_t v = (int)lane <= first ? (uint32_t)(d & LB_VALUE) : 0u;
#pragma unroll
    for(int o = 16; o; o >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, o);
    excl += v;
    if(first < 32)
      break;
    window -= 32;
  }
  if(lane == 0)
    lbStore(desc + bid, LB_INCLUSIVE | (unsigned long long)(excl + total));
  return excl;
}

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS   = 8;
constexpr int SCAN_TILE    = SCAN_THREADS * SCAN_ITEMS;

// out may alias in; out[n] = grand total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_lookback(const uint32_t* in, uint32_t* out, size_t n, unsigned long long* desc,
                                                                uint32_t* ticket, unsigned long long* stats)
{
  __shared__ uint32_t sm[33];
  __shared__ uint32_t sBid, sBase;
  if(threadIdx.x == 0)
    sBid = atomicAdd(ticket, 1u);
  __syncthreads();
  const uint32_t bid  = sBid;
  const size_t   base = (size_t)bid * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t       v[SCAN_ITEMS];
  uint32_t       acc = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; k++)
  {
    v[k] = base + k < n ? in[base + k] : 0u;
    acc += v[k];
  }
  uint32_t total;
  uint32_t ex = blockExclusiveScan(acc, sm, total);
  if(threadIdx.x < 32)
  {
    const uint32_t excl = lookbackExclusive(desc, bid, total, stats);
    if(threadIdx.x == 0)
      sBase = excl;
  }
  __syncthreads();
  ex += sBase;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; k++)
  {
    if(base + k < n)
      out[base + k] = ex;
    ex += v[k];
  }
  if(bid == gridDim.x - 1 && threadIdx.x == 0)
    out[n] = sBase + total;
}

static size_t scanBlocks(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// exclusive scan of in[0..n) into out[0..n], out[n] = total.  desc: scanBlocks(n) zeroed descriptors, ticket: one zeroed word
static int launchScan(const uint32_t* in, uint32_t* out, size_t n, unsigned long long* desc, uint32_t* ticket, unsigned long long* stats,
                      cudaStream_t s)
{
  const size_t nb = scanBlocks(n);
  if(nb == 0)
  {
    cudaMemsetAsync(out, 0, sizeof(uint32_t), s);
    return 0;
  }
  k_scan_lookback<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, out, n, desc, ticket, stats);
  return 1;
}

// ------------------------------------------------------------------------------------------------------------------
// stable LSD radix sort of (key, value) pairs on 8-bit digits; the element count is read from device memory
// ------------------------------------------------------------------------------------------------------------------
constexpr int SORT_THREADS = 256;
constexpr int SORT_ROUNDS  = 16;
constexpr int SORT_TILE    = SORT_THREADS * SORT_ROUNDS;

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ nPtr, int shift,
                                                            uint32_t* __restrict__ table, uint32_t nb)
{
  __shared__ uint32_t hist[256];
  const size_t        n = *nPtr;
  hist[threadIdx.x]     = 0;
  __syncthreads();
  const size_t base = (size_t)blockIdx.x * SORT_TILE;
  if(base < n)
  {
    // all the block's keys are requested before the first one is used (the kernel is nothing but load latency)
    uint32_t k[SORT_ROUNDS];
#pragma unroll
    for(int r = 0; r < SORT_ROUNDS; r++)
    {
      const size_t i = base + (size_t)r * SORT_THREADS + threadIdx.x;
      k[r]           = i < n ? keys[i] : 0xFFFFFFFFu;
    }
#pragma unroll
    for(int r = 0; r < SORT_ROUNDS; r++)
      if(base + (size_t)r * SORT_THREADS + threadIdx.x < n)
        atomicAdd(&hist[(k[r] >> shift) & 255u], 1u);
  }
  __syncthreads();
  table[(size_t)threadIdx.x * nb + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const uint32_t* __restrict__ keysIn, const uint32_t* __restrict__ valsIn,
                                                               uint32_t* __restrict__ keysOut, uint32_t* __restrict__ valsOut,
                                                               const uint32_t* __restrict__ nPtr, int shift,
                                                               const uint32_t* __restrict__ table, uint32_t nb)
{
  __shared__ uint32_t off[256];
  __shared__ uint32_t cnt[SORT_THREADS / 32][256];
  const size_t        n    = *nPtr;
  const size_t        base = (size_t)blockIdx.x * SORT_TILE;
  if(base >= n)
    return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  off[threadIdx.x] = table[(size_t)threadIdx.x * nb + blockIdx.x];
#pragma unroll
  for(int w = 0; w < SORT_THREADS / 32; w++)
    cnt[w][threadIdx.x] = 0;
  // all the block's pairs are requested up front: the rounds below are separated by barriers, and a load inside a round
  // would expose its full latency sixteen times
  uint32_t keyR[SORT_ROUNDS], valR[SORT_ROUNDS];
#pragma unroll
  for(int r = 0; r < SORT_ROUNDS; r++)
  {
    const size_t i = base + (size_t)r * SORT_THREADS + threadIdx.x;
    keyR[r]        = i < n ? keysIn[i] : 0u;
    valR[r]        = i < n ? valsIn[i] : 0u;
  }
  __syncthreads();
#pragma unroll
  for(int r = 0; r < SORT_ROUNDS; r++)
  {
    const size_t   i      = base + (size_t)r * SORT_THREADS + threadIdx.x;
    const bool     active = i < n;
    const uint32_t key = keyR[r], val = valR[r];
    const uint32_t d      = active ? ((key >> shift) & 255u) : 256u;
    const uint32_t peers  = __match_any_sync(0xffffffffu, d);
    const uint32_t rank   = __popc(peers & ((1u << lane) - 1u));
    if(active && rank == 0)
      cnt[warp][d] = __popc(peers);
    __syncthreads();
    if(active)
    {
      uint32_t prefix = 0;
      for(int w = 0; w < warp; w++)
        prefix += cnt[w][d];
      const uint32_t dst = off[d] + prefix + rank;
      keysOut[dst]       = key;
      valsOut[dst]       = val;
    }
    __syncthreads();
    {
      uint32_t total = 0;
#pragma unroll
      for(int w = 0; w < SORT_THREADS / 32; w++)
      {
        total += cnt[w][threadIdx.x];
        cnt[w][threadIdx.x] = 0;
      }
      off[threadIdx.x] += total;
    }
    __syncthreads();
  }
}

// tileStart[t] = index of the first pair whose key is >= t; tileStart[numTiles] = n
__global__ void __launch_bounds__(256) k_tile_ranges(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ nPtr, uint32_t numTiles,
                                                     uint32_t* __restrict__ tileStart)
{
  const uint32_t n = *nPtr;
  if(n == 0)
  {
    for(uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t <= numTiles; t += gridDim.x * blockDim.x)
      tileStart[t] = 0;
    return;
  }
  for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
  {
    const uint32_t k    = keys[i];
    const uint32_t prev = i == 0 ? 0u : keys[i - 1] + 1u;
    for(uint32_t t = prev; t <= k; t++)
      tileStart[t] = i;
    if(i == n - 1)
      for(uint32_t t = k + 1; t <= numTiles; t++)
        tileStart[t] = n;
  }
}

// Launch order of the raster CTAs: empty tiles, then heaviest tiles first, so that the light tiles fill the tail of the grid
// (longest-processing-time-first scheduling).  A counting sort on min(list length, 1022): a histogram pass, then every CTA
// scans the 1024 bins for itself and places its tiles with one atomic per tile; the order among tiles of equal weight is
// arbitrary, which is fine: tiles are independent, the order only affects scheduling.  (bins / cursors: zeroed with the
// rest of the binning's per-frame state.)
constexpr int ORDER_BINS = 1024;
__device__ __forceinline__ uint32_t orderBin(const uint32_t* __restrict__ tileStart, uint32_t t)
{
  // bin 0 = the empty tiles (their CTAs only write the clear colour: over in no time, and in split-frame mode their pixels
  // travel to the other bands while the heavy tiles are rasterised instead of in a burst at the end), bin 1 = the heaviest
  const uint32_t len = tileStart[t + 1] - tileStart[t];
  return len == 0u ? 0u : (uint32_t)ORDER_BINS - 1u - min(len, (uint32_t)ORDER_BINS - 2u);
}
__global__ void __launch_bounds__(256) k_tile_hist(const uint32_t* __restrict__ tileStart, uint32_t numTiles, uint32_t* __restrict__ bins)
{
  for(uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < numTiles; t += gridDim.x * blockDim.x)
    atomicAdd(&bins[orderBin(tileStart, t)], 1u);
}
__global__ void __launch_bounds__(1024) k_tile_order(const uint32_t* __restrict__ tileStart, uint32_t numTiles, const uint32_t* __restrict__ bins,
                                                     uint32_t* __restrict__ cursors, uint32_t* __restrict__ order)
{
  __shared__ uint32_t binStart[ORDER_BINS];
  __shared__ uint32_t scanSm[33];
  uint32_t            total;
  binStart[threadIdx.x] = blockExclusiveScan(bins[threadIdx.x], scanSm, total);
  __syncthreads();
  for(uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < numTiles; t += gridDim.x * blockDim.x)
  {
    const uint32_t b = orderBin(tileStart, t);
    order[binStart[b] + atomicAdd(&cursors[b], 1u)] = t;
  }
}

static size_t   sortBlocks(size_t pairCapacity) { return (pairCapacity + SORT_TILE - 1) / SORT_TILE; }
static uint32_t binBlocks(size_t triCount) { return (uint32_t)((triCount + BIN_BLOCK - 1) / BIN_BLOCK); }

// scratch of the sort: the [digit][block] histogram table (+ its grand total)
size_t binScratchWords(size_t /*triCount*/, size_t pairCapacity, size_t /*numTiles*/) { return 256 * sortBlocks(pairCapacity) + 2; }

// per-frame state of one binning, zeroed by one memset: [info: 4 words][tickets: 4 words][tile-order bins + cursors]
// [look-back descriptors of the two table scans][pair count of every block of triangles]
constexpr size_t LB_HEADER_BYTES = 32 + 2 * ORDER_BINS * sizeof(uint32_t);
size_t binLookbackBytes(size_t triCount, size_t pairCapacity)
{
  return LB_HEADER_BYTES + sizeof(unsigned long long) * (2 * scanBlocks(256 * sortBlocks(pairCapacity)) + 2) + sizeof(uint32_t) * ((size_t)binBlocks(triCount) + 2);
}

// Bins the triangles [firstTri, firstTri + triCount) of the index buffer.  On return (asynchronously) b.pairInfo[0] holds
// the number of pairs, b.pairVal[*sortedBuf] the triangle lists and b.tileStart the per-tile ranges.
int launchBin(const FrameParams& p, const BinBuffers& b, uint32_t firstTri, uint32_t triCount, bool cullBack, int* sortedBuf,
              cudaStream_t s)
{
  const uint32_t numTiles = (uint32_t)p.tilesX * p.tileRowsLocal;
  int            launches = 0;
  *sortedBuf              = 0;
  if(triCount == 0 || numTiles == 0)
  {
    cudaMemsetAsync(b.tileStart, 0, sizeof(uint32_t) * (numTiles + 1), s);
    cudaMemsetAsync(b.pairInfo, 0, sizeof(uint32_t) * 2, s);
    return 0;
  }
  // one memset: pair / clip counters, tickets, tile-order bins and the look-back descriptors of this binning
  cudaMemsetAsync(b.lb, 0, b.lbBytes, s);
  const uint32_t      nb       = (uint32_t)sortBlocks(b.pairCapacity);
  const size_t        tableLen = (size_t)256 * nb;
  uint32_t*           tickets  = b.pairInfo + 4;
  uint32_t*           orderBins = b.pairInfo + 8;
  unsigned long long* descScan[2];
  descScan[0]           = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(b.lb) + LB_HEADER_BYTES);
  descScan[1]           = descScan[0] + scanBlocks(tableLen) + 1;
  uint32_t* blockTotals = reinterpret_cast<uint32_t*>(descScan[1] + scanBlocks(tableLen) + 1);
  const uint32_t blocks = binBlocks(triCount);
  k_bin_count<<<blocks, 256, 0, s>>>(p, firstTri, triCount, cullBack ? 1 : 0, blockTotals);
  k_bin_emit<<<blocks, 256, 0, s>>>(p, firstTri, triCount, cullBack ? 1 : 0, blockTotals, b.pairKey[0], b.pairVal[0], (uint32_t)b.pairCapacity,
                                    b.pairInfo);
  launches += 2;
  int bits = 1;
  while((1u << bits) < numTiles)
    bits++;
  uint32_t* table = b.scratch;
  int       cur   = 0;
  for(int shift = 0, pass = 0; shift < bits; shift += 8, pass++)
  {
    k_sort_hist<<<nb, SORT_THREADS, 0, s>>>(b.pairKey[cur], b.pairInfo, shift, table, nb);
    launches += 1 + launchScan(table, table, tableLen, descScan[pass & 1], tickets + (pass & 1), p.stats, s);
    k_sort_scatter<<<nb, SORT_THREADS, 0, s>>>(b.pairKey[cur], b.pairVal[cur], b.pairKey[cur ^ 1], b.pairVal[cur ^ 1], b.pairInfo, shift,
                                               table, nb);
    launches++;
    cur ^= 1;
  }
  const int rblocks = (int)min((uint32_t)((b.pairCapacity + 255u) / 256u), 148u * 8u);
  k_tile_ranges<<<rblocks, 256, 0, s>>>(b.pairKey[cur], b.pairInfo, numTiles, b.tileStart);
  launches++;
  *sortedBuf = cur;
  k_tile_hist<<<(int)min((numTiles + 255u) / 256u, 148u * 4u), 256, 0, s>>>(b.tileStart, numTiles, orderBins);
  k_tile_order<<<(int)min((numTiles + 1023u) / 1024u, 148u), 1024, 0, s>>>(b.tileStart, numTiles, orderBins, orderBins + ORDER_BINS, b.tileOrder);
  launches += 2;
  return launches;
}

}  // namespace oit
