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
    const TVert t     = finishVertex(clip, viewz, hw, hh);
    p.tv[i] = t;
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
  ClipInput in;
  int       msaa, stripTileRows, bandCount, bandIndex, tilesX;
};
__device__ __forceinline__ BinView binView(const FrameParams& p)
{
  return BinView{clipInput(p), p.msaa, p.stripTileRows, p.bandCount, p.bandIndex, p.tilesX};
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
  uint32_t  n  = 0;
  const int nx = r.tx1 - r.tx0 + 1;
  for(int R = r.ty0; R <= r.ty1; R++)
    if(tileRowOwner(R, v.stripTileRows, v.bandCount) == v.bandIndex)
      n += nx;
  return n;
}
__device__ __forceinline__ uint32_t emitTiles(const BinView& v, const TileRange& r, uint32_t val, uint32_t o, uint32_t* __restrict__ keys,
                                              uint32_t* __restrict__ vals)
{
  for(int R = r.ty0; R <= r.ty1; R++)
    if(tileRowOwner(R, v.stripTileRows, v.bandCount) == v.bandIndex)
    {
      const uint32_t rowKey = (uint32_t)tileRowToLocal(R, v.stripTileRows, v.bandCount) * v.tilesX;
      for(int tx = r.tx0; tx <= r.tx1; tx++, o++)
      {
        keys[o] = rowKey + tx;
        vals[o] = val;  // emitted in primitive order; the stable sort keeps it (and the pieces of a primitive) per tile
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
        const float* aP    = v.in.verts + (size_t)ix[cv.i] * 10;
        const float* aQ    = v.in.verts + (size_t)ix[cv.j] * 10;
        for(int c = 0; c < 10; c++)
          ce.attr[m][c] = cv.i == cv.j ? aP[c] : __fmaf_rn(cv.t, __fsub_rn(aQ[c], aP[c]), aP[c]);
      }
      ce.pad0   = 0u;
      ce.pad[0] = ce.pad[1] = 0u;
      val                   = PAIR_CLIPPED | e;
    }
    else
      atomicAdd(overflow, 1ull);  // the host grows the table and renders the frame again
    o = emitTiles(v, r, val, o, keys, vals);
  }
}

__global__ void __launch_bounds__(256) k_bin_count(const FrameParams p, uint32_t firstTri, uint32_t triCount, int cullBack,
                                                   uint32_t* __restrict__ counts, uint32_t* __restrict__ pairInfo)
{
  unsigned long long nRejected = 0;
  if(blockIdx.x == 0 && threadIdx.x == 0)
    pairInfo[2] = 0u;  // clip entries handed out by k_bin_emit (which runs after this kernel)
  const BinView v = binView(p);
  for(uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < triCount; t += gridDim.x * blockDim.x)
  {
    const uint32_t tri = firstTri + t;
    const uint32_t i0 = p.indices[3 * (size_t)tri], i1 = p.indices[3 * (size_t)tri + 1], i2 = p.indices[3 * (size_t)tri + 2];
    const TVert    a = p.tv[i0], b = p.tv[i1], c = p.tv[i2];
    uint32_t       n = 0;
    if(a.x != INT32_MIN && b.x != INT32_MIN && c.x != INT32_MIN)
    {
      TileRange r;
      if(tvTileRange(v, a, b, c, cullBack != 0, r))
        n = ownedTiles(v, r);
    }
    else
    {
      bool rejected;
      n = countClipped(v, i0, i1, i2, cullBack != 0, &rejected);
      nRejected += rejected ? 1u : 0u;
    }
    counts[t] = n;
  }
  if(nRejected)
    atomicAdd(&p.stats[STAT_REJECTED], nRejected);
}

// offsets[t] = exclusive prefix of the counts, offsets[triCount] = number of pairs.  Pairs beyond `capacity` are dropped
// and the overflow flag is raised (the host then grows the buffers and renders the frame again).
__global__ void __launch_bounds__(256) k_bin_emit(const FrameParams p, uint32_t firstTri, uint32_t triCount, int cullBack,
                                                  const uint32_t* __restrict__ offsets, uint32_t* __restrict__ keys,
                                                  uint32_t* __restrict__ vals, uint32_t capacity, uint32_t* __restrict__ pairInfo)
{
  if(blockIdx.x == 0 && threadIdx.x == 0)
  {
    const uint32_t total = offsets[triCount];
    pairInfo[0]          = total > capacity ? 0u : total;  // pairs present (none when the frame has to be redone)
    pairInfo[1]          = total;                 // pairs wanted
    if(total > capacity)
      atomicAdd(&p.stats[STAT_OVERFLOW], 1ull);
  }
  const BinView v = binView(p);
  for(uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < triCount; t += gridDim.x * blockDim.x)
  {
    const uint32_t o0 = offsets[t], o1 = offsets[t + 1];
    if(o0 == o1 || o1 > capacity)
      continue;
    const uint32_t tri = firstTri + t;
    const uint32_t i0 = p.indices[3 * (size_t)tri], i1 = p.indices[3 * (size_t)tri + 1], i2 = p.indices[3 * (size_t)tri + 2];
    const TVert    a = p.tv[i0], b = p.tv[i1], c = p.tv[i2];
    if(a.x != INT32_MIN && b.x != INT32_MIN && c.x != INT32_MIN)
    {
      TileRange r;
      if(tvTileRange(v, a, b, c, cullBack != 0, r))
        emitTiles(v, r, tri, o0, keys, vals);
    }
    else
      emitClipped(v, i0, i1, i2, cullBack != 0, o0, keys, vals, p.clipEntries, p.clipCapacity, pairInfo + 2, p.stats + STAT_OVERFLOW);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// exclusive scan of uint32 (three kernels; the middle one is a single CTA walking the block sums)
// ------------------------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS   = 8;
constexpr int SCAN_TILE    = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ sums)
{
  __shared__ uint32_t sm[33];
  const size_t        base = (size_t)blockIdx.x * SCAN_TILE;
  uint32_t            acc  = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; k++)
  {
    const size_t i = base + (size_t)threadIdx.x * SCAN_ITEMS + k;
    acc += i < n ? in[i] : 0u;
  }
  uint32_t total;
  blockExclusiveScan(acc, sm, total);
  if(threadIdx.x == 0)
    sums[blockIdx.x] = total;
}
// in place: sums[i] <- exclusive prefix; sums[nb] <- grand total
__global__ void __launch_bounds__(1024) k_scan_top(uint32_t* sums, size_t nb)
{
  __shared__ uint32_t sm[33];
  uint32_t            carry = 0;
  for(size_t base = 0; base < nb; base += blockDim.x)
  {
    const size_t   i = base + threadIdx.x;
    const uint32_t v = i < nb ? sums[i] : 0u;
    uint32_t       total;
    const uint32_t ex = blockExclusiveScan(v, sm, total);
    if(i < nb)
      sums[i] = carry + ex;
    carry += total;
  }
  if(threadIdx.x == 0)
    sums[nb] = carry;
}
// out[i] = exclusive prefix of in (may alias); out[n] = grand total
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(const uint32_t* in, size_t n, const uint32_t* __restrict__ sums,
                                                             uint32_t* out, size_t nb)
{
  __shared__ uint32_t sm[33];
  const size_t        base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
  uint32_t            v[SCAN_ITEMS];
  uint32_t            acc = 0;
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; k++)
  {
    v[k] = base + k < n ? in[base + k] : 0u;
    acc += v[k];
  }
  uint32_t total;
  uint32_t ex = blockExclusiveScan(acc, sm, total) + sums[blockIdx.x];
#pragma unroll
  for(int k = 0; k < SCAN_ITEMS; k++)
  {
    if(base + k < n)
      out[base + k] = ex;
    ex += v[k];
  }
  if(blockIdx.x == 0 && threadIdx.x == 0)
    out[n] = sums[nb];
}

static size_t scanBlocks(size_t n) { return (n + SCAN_TILE - 1) / SCAN_TILE; }

// exclusive scan of in[0..n) into out[0..n], out[n] = total. scratch: scanBlocks(n) + 1 words
static int launchScan(const uint32_t* in, uint32_t* out, size_t n, uint32_t* scratch, cudaStream_t s)
{
  const size_t nb = scanBlocks(n);
  if(nb == 0)
  {
    cudaMemsetAsync(out, 0, sizeof(uint32_t), s);
    return 0;
  }
  k_scan_sums<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, scratch);
  k_scan_top<<<1, 1024, 0, s>>>(scratch, nb);
  k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(in, n, scratch, out, nb);
  return 3;
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
#pragma unroll 4
    for(int r = 0; r < SORT_ROUNDS; r++)
    {
      const size_t i = base + (size_t)r * SORT_THREADS + threadIdx.x;
      if(i < n)
        atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
    }
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
  __syncthreads();
  for(int r = 0; r < SORT_ROUNDS; r++)
  {
    const size_t   i      = base + (size_t)r * SORT_THREADS + threadIdx.x;
    const bool     active = i < n;
    const uint32_t key    = active ? keysIn[i] : 0u;
    const uint32_t val    = active ? valsIn[i] : 0u;
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

// Launch order of the raster CTAs: heaviest tiles first, so that the light tiles fill the tail of the grid
// (longest-processing-time-first scheduling).  A counting sort on min(list length, 1023) by ONE CTA; the order among
// tiles of equal weight is arbitrary, which is fine: tiles are independent, the order only affects scheduling.
constexpr int ORDER_BINS = 1024;
__global__ void __launch_bounds__(1024) k_tile_order(const uint32_t* __restrict__ tileStart, uint32_t numTiles, uint32_t* __restrict__ order)
{
  __shared__ uint32_t bins[ORDER_BINS];
  __shared__ uint32_t scanSm[33];
  const int           tid = threadIdx.x;
  bins[tid]               = 0;
  __syncthreads();
  for(uint32_t t = tid; t < numTiles; t += 1024)
    atomicAdd(&bins[ORDER_BINS - 1 - min(tileStart[t + 1] - tileStart[t], (uint32_t)ORDER_BINS - 1)], 1u);  // bin 0 = heaviest
  __syncthreads();
  const uint32_t mine = bins[tid];
  uint32_t       total;
  const uint32_t excl = blockExclusiveScan(mine, scanSm, total);
  bins[tid]           = excl;
  __syncthreads();
  for(uint32_t t = tid; t < numTiles; t += 1024)
    order[atomicAdd(&bins[ORDER_BINS - 1 - min(tileStart[t + 1] - tileStart[t], (uint32_t)ORDER_BINS - 1)], 1u)] = t;
}

size_t binScratchWords(size_t triCount, size_t pairCapacity, size_t /*numTiles*/)
{
  const size_t sortBlocks = (pairCapacity + SORT_TILE - 1) / SORT_TILE;
  const size_t table      = 256 * sortBlocks + 1;
  return scanBlocks(triCount) + 2 + table + scanBlocks(table) + 2;
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
  const int blocks = (int)min((triCount + 255u) / 256u, 148u * 16u);
  k_bin_count<<<blocks, 256, 0, s>>>(p, firstTri, triCount, cullBack ? 1 : 0, b.counts, b.pairInfo);
  launches += 1 + launchScan(b.counts, b.counts, triCount, b.scratch, s);
  k_bin_emit<<<blocks, 256, 0, s>>>(p, firstTri, triCount, cullBack ? 1 : 0, b.counts, b.pairKey[0], b.pairVal[0],
                                    (uint32_t)b.pairCapacity, b.pairInfo);
  launches++;
  int bits = 1;
  while((1u << bits) < numTiles)
    bits++;
  const uint32_t nb        = (uint32_t)((b.pairCapacity + SORT_TILE - 1) / SORT_TILE);
  uint32_t*      table     = b.scratch + scanBlocks(triCount) + 2;
  uint32_t*      tableScan = table + (size_t)256 * nb + 1;
  int            cur       = 0;
  for(int shift = 0; shift < bits; shift += 8)
  {
    k_sort_hist<<<nb, SORT_THREADS, 0, s>>>(b.pairKey[cur], b.pairInfo, shift, table, nb);
    launches += 1 + launchScan(table, table, (size_t)256 * nb, tableScan, s);
    k_sort_scatter<<<nb, SORT_THREADS, 0, s>>>(b.pairKey[cur], b.pairVal[cur], b.pairKey[cur ^ 1], b.pairVal[cur ^ 1], b.pairInfo, shift,
                                               table, nb);
    launches++;
    cur ^= 1;
  }
  const int rblocks = (int)min((uint32_t)((b.pairCapacity + 255u) / 256u), 148u * 8u);
  k_tile_ranges<<<rblocks, 256, 0, s>>>(b.pairKey[cur], b.pairInfo, numTiles, b.tileStart);
  launches++;
  *sortedBuf = cur;
  k_tile_order<<<1, 1024, 0, s>>>(b.tileStart, numTiles, b.tileOrder);
  launches++;
  return launches;
}

}  // namespace oit
