// oit_raster.cu -- tile-ordered software rasteriser + the per-technique fragment ("colour pass") programs.
//
// Replaces, for one geometry pass of the reference (vkCmdDrawIndexed of the transparent or opaque range):
//   fixed-function raster / early per-sample depth test / post-depth coverage   main.cpp:504-532, oitColorDepthDefines.glsl:36-37
//   shading()                                                                    shaderCommon.glsl:36-56
//   K2  oitSimple.frag.glsl:50-91        K4  oitLinkedList.frag.glsl:51-85    K6/K7 oitLoop.frag.glsl:57-100,120-173
//   K9  oitLoop64.frag.glsl:65-141       K11 oitSpinlock.frag.glsl:49-129     K13  oitInterlock.frag.glsl:90-152
//   K15 oitWeighted.frag.glsl:53-79      K1  opaque.frag.glsl:30-34
//   ROP blending in primitive order                                              main.cpp:540-592
//
// Execution model: one CTA per 16x16 screen tile.  The CTA walks the tile's triangle list (primitive order) in
// chunks of 256 triangles: thread t stages triangle t of the chunk (edge setup, bounding box clipped to the tile),
// a CTA-wide scan turns the boxes into a dense list of (triangle, pixel) work items, and the threads then take the
// items 256 at a time.  Two items of the same pixel inside one round are serialised in item order through a
// per-pixel owner word in shared memory (atomicMin arbitration), so every pixel sees its fragments in PRIMITIVE
// ORDER -- which is what the hardware ROP guarantees for the tail blend and what the ordered interlock asks for.
// Because a tile is owned by one CTA, the A-buffer slice, aux words and colour samples of the tile stay in that
// SM's L1/L2 for the whole pass.
#include <cooperative_groups.h>

#include "oit_device.cuh"

namespace cg = cooperative_groups;

namespace oit {

struct TriSlot
{
  int32_t  x[3], y[3];    // snapped vertex positions, re-ordered so that area2 > 0
  float    z0, dz1, dz2;  // screen-linear depth plane through vertex 0
  float    iw[3];
  uint32_t vidx[3];
  float    farea;         // (float)area2
  uint32_t box;           // bx0 | by0 << 4 | (bw-1) << 8 | (bh-1) << 12 | bias bits << 16 | zSafe << 19
  uint32_t rcpW;          // ceil(65536 / bw)
};

struct FragCtx
{
  const FrameParams& p;
  const SrgbTables&  t;
  uint32_t           nFrag, nStored, nTail, nOpaque;
};

__device__ __forceinline__ uint32_t ldcg32(const uint32_t* a) { return __ldcg(a); }
__device__ __forceinline__ unsigned long long ldcg64(const unsigned long long* a) { return __ldcg(a); }

// ---- varyings + shading -------------------------------------------------------------------------------------------
struct Bary
{
  float l0, l1, l2;
};
__device__ __forceinline__ Bary makeBary(long long e1, long long e2, float farea)
{
  Bary b;
  b.l1 = __fdiv_rn(__ll2float_rn(e1), farea);
  b.l2 = __fdiv_rn(__ll2float_rn(e2), farea);
  b.l0 = __fsub_rn(__fsub_rn(1.0f, b.l1), b.l2);
  return b;
}
__device__ __forceinline__ float depthAt(const TriSlot& s, const Bary& b)
{
  return clamp01(__fmaf_rn(b.l2, s.dz2, __fmaf_rn(b.l1, s.dz1, s.z0)));
}

// Interpolants (shaderCommon.glsl:25-31) perspective-correct at `b`, then shading() (shaderCommon.glsl:36-56)
template <bool NEED_VIEWZ>
__device__ __forceinline__ Color4 shadeAt(const FrameParams& p, const TriSlot& s, const Bary& b, float& viewz)
{
  const float q0 = __fmul_rn(b.l0, s.iw[0]), q1 = __fmul_rn(b.l1, s.iw[1]), q2 = __fmul_rn(b.l2, s.iw[2]);
  const float rden = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(q0, q1), q2));
  const float* a0 = p.verts + (size_t)s.vidx[0] * 10;
  const float* a1 = p.verts + (size_t)s.vidx[1] * 10;
  const float* a2 = p.verts + (size_t)s.vidx[2] * 10;
  float        v[7];
#pragma unroll
  for(int k = 0; k < 7; k++)
    v[k] = __fmul_rn(__fmaf_rn(q2, __ldg(a2 + 3 + k), __fmaf_rn(q1, __ldg(a1 + 3 + k), __fmul_rn(q0, __ldg(a0 + 3 + k)))), rden);
  if(NEED_VIEWZ)
    viewz = __fmul_rn(__fmaf_rn(q2, p.tv[s.vidx[2]].viewz, __fmaf_rn(q1, p.tv[s.vidx[1]].viewz, __fmul_rn(q0, p.tv[s.vidx[0]].viewz))), rden);
  const float LX = -0.40824829046386301637f, LY = 0.81649658092772603273f, LZ = 0.40824829046386301637f;
  const float len2 = __fmaf_rn(v[2], v[2], __fmaf_rn(v[1], v[1], __fmul_rn(v[0], v[0])));
  float       nx = 0.f, ny = 0.f, nz = 0.f;
  if(len2 > 0.f)
  {
    const float inv = __fdiv_rn(1.0f, __fsqrt_rn(len2));
    nx              = __fmul_rn(v[0], inv);
    ny              = __fmul_rn(v[1], inv);
    nz              = __fmul_rn(v[2], inv);
  }
  const float d      = __fmaf_rn(nz, LZ, __fmaf_rn(ny, LY, __fmul_rn(nx, LX)));
  const float warmth = __fmaf_rn(d, 0.5f, 0.5f);
  const float om     = __fsub_rn(1.0f, warmth);
  Color4      c;
  c.r = __fmul_rn(v[3], __fmaf_rn(0.0f, om, warmth));
  c.g = __fmul_rn(v[4], __fmaf_rn(0.25f, om, warmth));
  c.b = __fmul_rn(v[5], __fmaf_rn(0.75f, om, warmth));
  c.a = clamp01(__fmaf_rn(v[6], p.alphaWidth, p.alphaMin));
  return c;
}

// ---- fragment programs ----------------------------------------------------------------------------------------------
// All return the colour handed to the ROP (premultiplied; zero = no-op).  x, yl: pixel (yl = row inside this band's
// buffers); pix = yl * W + x; ai = aux index of (sampleID, pixel).

// K2 oitSimple.frag.glsl:50-91
__device__ __forceinline__ Color4 fragSimple(FragCtx& c, size_t pix, size_t ai, uint32_t sampleID, uint32_t mask, const Color4& rgba, float z)
{
  const FrameParams& p        = c.p;
  const size_t       viewSize = (size_t)p.W * p.localH;
  const size_t       listPos  = viewSize * p.L * sampleID + pix;
  const uint32_t     old      = atomicAdd(&p.aux[ai], 1u);
  if(old < (uint32_t)p.L)
  {
    const uint32_t packed = packColor(c.t, rgba);
    if(p.coverage)
      reinterpret_cast<uint4*>(p.abuf)[listPos + (size_t)old * viewSize] = make_uint4(packed, __float_as_uint(z), mask, 0u);
    else
      reinterpret_cast<uint2*>(p.abuf)[listPos + (size_t)old * viewSize] = make_uint2(packed, __float_as_uint(z));
    c.nStored++;
    return zeroColor();
  }
  if(p.tailBlend)
  {
    c.nTail++;
    return premultiply(rgba);
  }
  return zeroColor();
}

// K4 oitLinkedList.frag.glsl:51-85 -- the single-address counter is bumped once per converged warp group
__device__ __forceinline__ Color4 fragLinkedList(FragCtx& c, size_t ai, uint32_t mask, const Color4& rgba, float z)
{
  const FrameParams& p = c.p;
  uint32_t           newOffset;
  {
    cg::coalesced_group g = cg::coalesced_threads();
    uint32_t            base = 0;
    if(g.thread_rank() == 0)
      base = atomicAdd(p.counter, g.size());
    newOffset = g.shfl(base, 0) + g.thread_rank() + 1u;
  }
  if(newOffset >= p.capacity)
  {
    if(p.tailBlend)
    {
      c.nTail++;
      return premultiply(rgba);
    }
    return zeroColor();
  }
  const uint32_t oldOffset = atomicExch(&p.aux[ai], newOffset);
  reinterpret_cast<uint4*>(p.abuf)[newOffset] =
      make_uint4(packColor(c.t, rgba), __float_as_uint(z), p.coverage ? mask : 0u, oldOffset);
  c.nStored++;
  return zeroColor();
}

// K6 oitLoop.frag.glsl:57-100 (depth pass)
__device__ __forceinline__ void fragLoopDepth(FragCtx& c, size_t pix, uint32_t sampleID, float z)
{
  const FrameParams& p        = c.p;
  const size_t       viewSize = (size_t)p.W * p.localH;
  uint32_t*          list     = p.abuf + viewSize * p.L * 2 * sampleID + pix;
  uint32_t           zcur     = __float_as_uint(z);
  int                i        = 0;
  uint32_t           pretest  = ldcg32(list + (size_t)(p.L - 1) * viewSize);
  if(zcur > pretest)
    return;
  pretest = ldcg32(list + (size_t)(p.L / 2) * viewSize);
  if(zcur > pretest)
    i = p.L / 2;
  for(; i < p.L; i++)
  {
    const uint32_t ztest = atomicMin(list + (size_t)i * viewSize, zcur);
    if(ztest == 0xFFFFFFFFu || ztest == zcur)
      break;
    zcur = max(ztest, zcur);
  }
}
// K7 oitLoop.frag.glsl:120-173 (colour pass)
__device__ __forceinline__ Color4 fragLoopColor(FragCtx& c, size_t pix, uint32_t sampleID, const Color4& rgba, float z)
{
  const FrameParams& p        = c.p;
  const size_t       viewSize = (size_t)p.W * p.localH;
  uint32_t*          list     = p.abuf + viewSize * p.L * 2 * sampleID + pix;
  const uint32_t     zcur     = __float_as_uint(z);
  if(list[(size_t)(p.L - 1) * viewSize] < zcur)
  {
    if(p.tailBlend)
    {
      c.nTail++;
      return premultiply(rgba);
    }
    return zeroColor();
  }
  int start = 0, end = p.L - 1;
  while(start < end)
  {
    const int      mid   = (start + end) / 2;
    const uint32_t ztest = list[(size_t)mid * viewSize];
    if(ztest < zcur)
      start = mid + 1;
    else
      end = mid;
  }
  list[(size_t)(p.L + start) * viewSize] = packColor(c.t, rgba);
  c.nStored++;
  return zeroColor();
}

// K9 oitLoop64.frag.glsl:65-141: key = depth << 32 | rgba8, cascade of 64-bit atomicMin
__device__ __forceinline__ Color4 fragLoop64(FragCtx& c, size_t pix, uint32_t sampleID, const Color4& rgba, float z)
{
  const FrameParams&  p        = c.p;
  const size_t        viewSize = (size_t)p.W * p.localH;
  unsigned long long* list     = reinterpret_cast<unsigned long long*>(p.abuf) + viewSize * p.L * sampleID + pix;
  unsigned long long  zcur     = ((unsigned long long)__float_as_uint(z) << 32) | packColor(c.t, rgba);
  int                 i        = 0;
  bool                canInsert = true;
  unsigned long long  pretest   = ldcg64(list + (size_t)(p.L - 1) * viewSize);
  if(zcur > pretest)
    canInsert = false;
  else
  {
    pretest = ldcg64(list + (size_t)(p.L / 2) * viewSize);
    if(zcur > pretest)
      i = p.L / 2;
  }
  bool evict = true;
  if(canInsert)
  {
    for(; i < p.L; i++)
    {
      const unsigned long long ztest = atomicMin(list + (size_t)i * viewSize, zcur);
      if(ztest == ~0ull)
      {
        evict = false;
        break;
      }
      zcur = ztest > zcur ? ztest : zcur;
    }
  }
  if(!evict)
  {
    c.nStored++;
    return zeroColor();
  }
  if(p.tailBlend)
  {
    c.nTail++;
    return premultiply(unpackColor(c.t, (uint32_t)(zcur & 0xFFFFFFFFull)));
  }
  return zeroColor();
}

// the critical section shared by K11 (oitSpinlock.frag.glsl:86-118) and K13 (oitInterlock.frag.glsl:110-145).
// Runs with the pixel exclusively owned (tile-ordered arbitration), so plain loads / stores are sufficient.
__device__ __forceinline__ bool lockCriticalSection(FragCtx& c, size_t pix, size_t ai, uint32_t sampleID, uint32_t mask, uint32_t packed,
                                                    uint32_t zbits, Color4& color)
{
  const FrameParams& p        = c.p;
  const size_t       viewSize = (size_t)p.W * p.localH;
  const size_t       listPos  = viewSize * p.L * sampleID + pix;
  const uint32_t     oldCounter = p.aux[ai];
  p.aux[ai]                     = oldCounter + 1u;
  if(oldCounter < (uint32_t)p.L)
  {
    if(p.coverage)
      reinterpret_cast<uint4*>(p.abuf)[listPos + (size_t)oldCounter * viewSize] = make_uint4(packed, zbits, mask, 0u);
    else
      reinterpret_cast<uint2*>(p.abuf)[listPos + (size_t)oldCounter * viewSize] = make_uint2(packed, zbits);
    color = zeroColor();
    return true;
  }
  int      furthest = 0;
  uint32_t maxDepth = 0;
  for(int i = 0; i < p.L; i++)
  {
    const size_t   e         = listPos + (size_t)i * viewSize;
    const uint32_t testDepth = p.coverage ? p.abuf[e * 4 + 1] : p.abuf[e * 2 + 1];
    if(testDepth > maxDepth)
    {
      maxDepth = testDepth;
      furthest = i;
    }
  }
  if(maxDepth > zbits)
  {
    const size_t e = listPos + (size_t)furthest * viewSize;
    if(p.coverage)
    {
      color                               = unpackColor(c.t, p.abuf[e * 4]);
      reinterpret_cast<uint4*>(p.abuf)[e] = make_uint4(packed, zbits, mask, 0u);
    }
    else
    {
      color                               = unpackColor(c.t, p.abuf[e * 2]);
      reinterpret_cast<uint2*>(p.abuf)[e] = make_uint2(packed, zbits);
    }
    p.adepth[ai] = maxDepth;
    return true;
  }
  return false;
}

// K11 oitSpinlock.frag.glsl:49-129 (keeps the reference's lock protocol: exchange-acquire, exchange-release, and the
// while(!done) shape that is deadlock-free under independent thread scheduling) and K13 oitInterlock.frag.glsl:90-152
template <bool SPIN>
__device__ __forceinline__ Color4 fragLock(FragCtx& c, size_t pix, size_t ai, uint32_t sampleID, uint32_t mask, const Color4& rgba, float z)
{
  const FrameParams& p      = c.p;
  const uint32_t     zbits  = __float_as_uint(z);
  const uint32_t     packed = packColor(c.t, rgba);
  Color4             color  = rgba;
  bool               stored = false;
  if(SPIN)
  {
    const uint32_t oldDepth = ldcg32(&p.adepth[ai]);  // racy-but-conservative early-out outside the lock (:67-68)
    if(zbits <= oldDepth)
    {
      bool done = mask == 0;
      while(!done)
      {
        const uint32_t old = atomicExch(&p.spin[ai], 1u);
        if(old == 0u)
        {
          stored = lockCriticalSection(c, pix, ai, sampleID, mask, packed, zbits, color);
          __threadfence();
          atomicExch(&p.spin[ai], 0u);
          done = true;
        }
      }
    }
  }
  else
  {
    // beginInvocationInterlock .. endInvocationInterlock: the arbitration loop of the caller IS the ordered interlock
    if(zbits <= p.adepth[ai])
      stored = lockCriticalSection(c, pix, ai, sampleID, mask, packed, zbits, color);
  }
  if(stored)
    c.nStored++;
  if(!p.tailBlend)
    return zeroColor();  // outColor never written in the reference (oitSpinlock.frag.glsl:126-128): defined as 0
  const Color4 out = premultiply(color);
  if(!isZero(out))
    c.nTail++;
  return out;
}

// K15 oitWeighted.frag.glsl:53-79 + BlendMode::WEIGHTED_COLOR into RGBA16F / R16F (main.cpp:559-575)
template <int S>
__device__ __forceinline__ void fragWeighted(FragCtx& c, size_t pix, uint32_t mask, const Color4& rgba, float viewz)
{
  const FrameParams& p   = c.p;
  const Color4       col = premultiply(rgba);
  const float        depthZ = __fmul_rn(-viewz, 10.0f);
  const float        x      = __fdiv_rn(depthZ, 200.0f);
  const float        x2     = __fmul_rn(x, x);
  const float        x4     = __fmul_rn(x2, x2);
  float              distWeight = __fdiv_rn(0.03f, __fadd_rn(1e-5f, x4));
  distWeight                    = distWeight < 1e-2f ? 1e-2f : (distWeight > 3e3f ? 3e3f : distWeight);
  const float mx     = fmaxf(fmaxf(col.r, col.g), fmaxf(col.b, col.a));
  float       aw     = fminf(1.0f, __fmaf_rn(mx, 40.0f, 0.01f));
  aw                 = __fmul_rn(aw, aw);
  const float weight = __fmul_rn(aw, distWeight);
  const float om     = __fsub_rn(1.0f, col.a);
  const float src[4] = {__fmul_rn(col.r, weight), __fmul_rn(col.g, weight), __fmul_rn(col.b, weight), __fmul_rn(col.a, weight)};
  uint16_t*   acc    = p.wacc + pix * S * 4;
  uint16_t*   rev    = p.wrev + pix * S;
#pragma unroll
  for(int s = 0; s < S; s++)
    if(mask & (1u << s))
    {
      ushort4 a = reinterpret_cast<ushort4*>(acc)[s];
      a.x       = f2h(__fadd_rn(h2f(a.x), src[0]));
      a.y       = f2h(__fadd_rn(h2f(a.y), src[1]));
      a.z       = f2h(__fadd_rn(h2f(a.z), src[2]));
      a.w       = f2h(__fadd_rn(h2f(a.w), src[3]));
      reinterpret_cast<ushort4*>(acc)[s] = a;
      rev[s]                             = f2h(__fmul_rn(h2f(rev[s]), om));
    }
  c.nStored++;
}

template <int S>
__device__ __forceinline__ void ropSamples(const FragCtx& c, size_t pix, uint32_t mask, const Color4& src)
{
  if(isZero(src))
    return;  // identity blend: encode(decode(v)) == v for every 8-bit v
  uint32_t* px = c.p.color + pix * S;
#pragma unroll
  for(int s = 0; s < S; s++)
    if(mask & (1u << s))
      px[s] = ropPremult(c.t, px[s], src);
}

// one colour-pass invocation + its ROP write
template <int PASS, int S>
__device__ __forceinline__ void invoke(FragCtx& c, int x, int yl, uint32_t sampleID, uint32_t mask, const Color4& rgba, float z, float viewz)
{
  const FrameParams& p   = c.p;
  const size_t       pix = (size_t)yl * p.W + x;
  const size_t       ai  = ((size_t)sampleID * p.localH + yl) * p.W + x;
  if(PASS == PASS_LOOP_DEPTH)
  {
    fragLoopDepth(c, pix, sampleID, z);
    return;
  }
  c.nFrag++;
  Color4 out = zeroColor();
  switch(PASS)
  {
    case PASS_SIMPLE: out = fragSimple(c, pix, ai, sampleID, mask, rgba, z); break;
    case PASS_LINKEDLIST: out = fragLinkedList(c, ai, mask, rgba, z); break;
    case PASS_LOOP_COLOR: out = fragLoopColor(c, pix, sampleID, rgba, z); break;
    case PASS_LOOP64: out = fragLoop64(c, pix, sampleID, rgba, z); break;
    case PASS_SPINLOCK: out = fragLock<true>(c, pix, ai, sampleID, mask, rgba, z); break;
    case PASS_INTERLOCK: out = fragLock<false>(c, pix, ai, sampleID, mask, rgba, z); break;
    case PASS_WEIGHTED: fragWeighted<S>(c, pix, mask, rgba, viewz); return;
  }
  ropSamples<S>(c, pix, mask, out);
}

// ---- the tile kernel ------------------------------------------------------------------------------------------------
template <int S>
__device__ __forceinline__ void sampleOffset(int s, int& sx, int& sy)
{
  samplePos(S, s, sx, sy);
}

template <int PASS, int S, bool SSHADE>
__global__ void __launch_bounds__(RASTER_THREADS) k_raster(const FrameParams p)
{
  __shared__ SrgbTables tabs;
  __shared__ TriSlot    slots[RASTER_THREADS];
  __shared__ uint32_t   itemStart[RASTER_THREADS];
  __shared__ uint32_t   owner[TILE_PIX];
  __shared__ uint32_t   scanSm[33];

  const int      tid  = threadIdx.x;
  const uint32_t tile = blockIdx.x;
  const uint32_t listBegin = p.tileStart[tile], listEnd = p.tileStart[tile + 1];
  if(listBegin == listEnd)
    return;
  const int rl = tile / p.tilesX, tx = tile - rl * p.tilesX;
  const int R  = tileRowToGlobal(rl, p.stripTileRows, p.bandCount, p.bandIndex);
  const int tileX0 = tx * TILE_W, tileY0 = R * TILE_H;  // global pixel origin of the tile
  const int yLocal0 = rl * TILE_H;                      // the same row inside this band's buffers

  loadTables(tabs, p.tables);
  owner[tid] = 0xFFFFFFFFu;
  FragCtx ctx{p, tabs, 0, 0, 0, 0};
  const int lo = S == 1 ? 128 : (S == 4 ? 32 : 16), hi = 256 - lo;
  __syncthreads();

  for(uint32_t base = listBegin; base < listEnd; base += RASTER_THREADS)
  {
    // ---- stage one triangle per thread ----------------------------------------------------------------------------
    uint32_t nItems = 0;
    if(base + tid < listEnd)
    {
      const uint32_t tri = p.pairTri[base + tid];
      uint32_t       ix[3] = {p.indices[3 * (size_t)tri], p.indices[3 * (size_t)tri + 1], p.indices[3 * (size_t)tri + 2]};
      TVert          v0 = p.tv[ix[0]], v1 = p.tv[ix[1]], v2 = p.tv[ix[2]];
      long long      area2 = (long long)(v1.x - v0.x) * (v2.y - v0.y) - (long long)(v2.x - v0.x) * (v1.y - v0.y);
      if(area2 < 0)
      {
        const TVert tv = v1;
        v1             = v2;
        v2             = tv;
        const uint32_t ti = ix[1];
        ix[1]             = ix[2];
        ix[2]             = ti;
        area2             = -area2;
      }
      TriSlot s;
      s.x[0] = v0.x; s.x[1] = v1.x; s.x[2] = v2.x;
      s.y[0] = v0.y; s.y[1] = v1.y; s.y[2] = v2.y;
      s.z0   = v0.z;
      s.dz1  = __fsub_rn(v1.z, v0.z);
      s.dz2  = __fsub_rn(v2.z, v0.z);
      s.iw[0] = v0.invw; s.iw[1] = v1.invw; s.iw[2] = v2.invw;
      s.vidx[0] = ix[0]; s.vidx[1] = ix[1]; s.vidx[2] = ix[2];
      s.farea = __ll2float_rn(area2);
      const int minx = min(v0.x, min(v1.x, v2.x)), maxx = max(v0.x, max(v1.x, v2.x));
      const int miny = min(v0.y, min(v1.y, v2.y)), maxy = max(v0.y, max(v1.y, v2.y));
      const int px0 = max((minx - hi + 255) >> 8, tileX0), px1 = min((maxx - lo) >> 8, min(tileX0 + TILE_W, p.W) - 1);
      const int py0 = max((miny - hi + 255) >> 8, tileY0), py1 = min((maxy - lo) >> 8, min(tileY0 + TILE_H, p.H) - 1);
      uint32_t  biasBits = 0;
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        const int a = (k + 1) % 3, b = (k + 2) % 3;
        const int dx = s.x[b] - s.x[a], dy = s.y[b] - s.y[a];
        const bool topLeft = (dy == 0 && dx > 0) || dy < 0;  // Vulkan / D3D top-left rule, y down, area2 > 0
        biasBits |= (topLeft ? 0u : 1u) << k;
      }
      const bool zSafe = fmaxf(v0.z, fmaxf(v1.z, v2.z)) < 0.9999f;
      s.box = 0;
      s.rcpW = 0;
      if(px0 <= px1 && py0 <= py1)
      {
        const int bw = px1 - px0 + 1, bh = py1 - py0 + 1;
        nItems       = bw * bh;
        s.box  = (uint32_t)(px0 - tileX0) | ((uint32_t)(py0 - tileY0) << 4) | ((uint32_t)(bw - 1) << 8) | ((uint32_t)(bh - 1) << 12)
                | (biasBits << 16) | ((zSafe ? 1u : 0u) << 19);
        s.rcpW = (65535u + bw) / bw;
      }
      slots[tid] = s;
    }
    uint32_t       total;
    const uint32_t ex = blockExclusiveScan(nItems, scanSm, total);
    itemStart[tid]    = ex;
    __syncthreads();

    // ---- consume the work items 256 at a time -------------------------------------------------------------------
    for(uint32_t k0 = 0; k0 < total; k0 += RASTER_THREADS)
    {
      const uint32_t k = k0 + tid;
      uint32_t       mask = 0;
      int            slot = 0, lx = 0, ly = 0;
      if(k < total)
      {
#pragma unroll
        for(int step = RASTER_THREADS / 2; step; step >>= 1)
          if(itemStart[slot + step] <= k)
            slot += step;
        const TriSlot& s     = slots[slot];
        const uint32_t local = k - itemStart[slot];
        const uint32_t row   = (local * s.rcpW) >> 16;
        const uint32_t bw    = ((s.box >> 8) & 15u) + 1u;
        lx                   = (int)((s.box & 15u) + (local - row * bw));
        ly                   = (int)(((s.box >> 4) & 15u) + row);
        // ---- coverage: three int64 edge functions per sample, top-left rule --------------------------------------
        const long long ox = (long long)(tileX0 + lx) << 8, oy = (long long)(tileY0 + ly) << 8;
        long long       e[3];
        int             dxs[3], dys[3];
#pragma unroll
        for(int q = 0; q < 3; q++)
        {
          const int a = (q + 1) % 3, b = (q + 2) % 3;
          dxs[q]      = s.x[b] - s.x[a];
          dys[q]      = s.y[b] - s.y[a];
          e[q]        = (long long)dxs[q] * (oy - s.y[a]) - (long long)dys[q] * (ox - s.x[a]) - (long long)((s.box >> (16 + q)) & 1u);
        }
        const bool   zSafe = (s.box >> 19) & 1u;
        const float* dpx   = p.depth ? p.depth + ((size_t)(yLocal0 + ly) * p.W + tileX0 + lx) * S : nullptr;
#pragma unroll
        for(int sI = 0; sI < S; sI++)
        {
          int sx, sy;
          sampleOffset<S>(sI, sx, sy);
          const long long e0 = e[0] + (long long)dxs[0] * sy - (long long)dys[0] * sx;
          const long long e1 = e[1] + (long long)dxs[1] * sy - (long long)dys[1] * sx;
          const long long e2 = e[2] + (long long)dxs[2] * sy - (long long)dys[2] * sx;
          if((e0 | e1 | e2) >= 0)
          {
            bool pass = true;
            if(dpx != nullptr || !zSafe)
            {
              // early per-sample depth test, VK_COMPARE_OP_LESS (main.cpp:530-532)
              const Bary  b  = makeBary(e1 + ((s.box >> 17) & 1u), e2 + ((s.box >> 18) & 1u), s.farea);
              const float zs = depthAt(s, b);
              pass           = zs < (dpx ? dpx[sI] : 1.0f);
            }
            if(pass)
              mask |= 1u << sI;
          }
        }
      }
      // ---- primitive-ordered execution: lowest item index wins each pixel -------------------------------------------
      const int pl      = ly * TILE_W + lx;
      bool      pending = mask != 0;
      while(__syncthreads_or(pending))
      {
        if(pending)
          atomicMin(&owner[pl], (uint32_t)tid);
        __syncthreads();
        if(pending && owner[pl] == (uint32_t)tid)
        {
          owner[pl] = 0xFFFFFFFFu;
          pending   = false;
          const TriSlot&  s  = slots[slot];
          const int       gx = tileX0 + lx, yl = yLocal0 + ly;
          const long long ox = (long long)gx << 8, oy = (long long)(tileY0 + ly) << 8;
          // unbiased edge functions 1 and 2 at the pixel origin
          const int       dx1 = s.x[0] - s.x[2], dy1 = s.y[0] - s.y[2], dx2 = s.x[1] - s.x[0], dy2 = s.y[1] - s.y[0];
          const long long e1o = (long long)dx1 * (oy - s.y[2]) - (long long)dy1 * (ox - s.x[2]);
          const long long e2o = (long long)dx2 * (oy - s.y[0]) - (long long)dy2 * (ox - s.x[0]);
          if(PASS == PASS_OPAQUE)
          {
            // opaque.frag.glsl:30-34, BlendMode::NONE with depth write (main.cpp:541-546); shaded at the pixel centre
            float        vz;
            const Bary   bc = makeBary(e1o + (long long)dx1 * 128 - (long long)dy1 * 128, e2o + (long long)dx2 * 128 - (long long)dy2 * 128, s.farea);
            Color4       g  = shadeAt<false>(p, s, bc, vz);
            g.a             = 1.0f;
            const uint32_t enc = encodeDst(tabs, g);
            const size_t   pix = (size_t)yl * p.W + gx;
            bool           wrote = false;
#pragma unroll
            for(int sI = 0; sI < S; sI++)
              if(mask & (1u << sI))
              {
                int sx, sy;
                sampleOffset<S>(sI, sx, sy);
                const Bary  b  = makeBary(e1o + (long long)dx1 * sy - (long long)dy1 * sx, e2o + (long long)dx2 * sy - (long long)dy2 * sx, s.farea);
                const float zs = depthAt(s, b);
                // the mask was computed before this pixel's earlier fragments of the same round ran: test again
                if(zs < p.depth[pix * S + sI])
                {
                  p.depth[pix * S + sI] = zs;
                  p.color[pix * S + sI] = enc;
                  wrote                 = true;
                }
              }
            ctx.nOpaque += wrote ? 1u : 0u;
          }
          else if(SSHADE && PASS != PASS_WEIGHTED)
          {
            // sample shading: every covered sample is its own invocation at the sample position
#pragma unroll
            for(int sI = 0; sI < S; sI++)
              if(mask & (1u << sI))
              {
                int sx, sy;
                sampleOffset<S>(sI, sx, sy);
                const Bary   b = makeBary(e1o + (long long)dx1 * sy - (long long)dy1 * sx, e2o + (long long)dx2 * sy - (long long)dy2 * sx, s.farea);
                float        vz = 0.f;
                const Color4 rgba = shadeAt<false>(p, s, b, vz);
                invoke<PASS, S>(ctx, gx, yl, (uint32_t)sI, 1u << sI, rgba, depthAt(s, b), vz);
              }
          }
          else
          {
            // one invocation per pixel, varyings and gl_FragCoord.z at the pixel centre
            const Bary   bc = makeBary(e1o + (long long)dx1 * 128 - (long long)dy1 * 128, e2o + (long long)dx2 * 128 - (long long)dy2 * 128, s.farea);
            float        vz = 0.f;
            const Color4 rgba = shadeAt<PASS == PASS_WEIGHTED>(p, s, bc, vz);
            invoke<PASS, S>(ctx, gx, yl, 0u, mask, rgba, depthAt(s, bc), vz);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- statistics ----------------------------------------------------------------------------------------------------
  uint32_t vals[4] = {ctx.nFrag, ctx.nStored, ctx.nTail, ctx.nOpaque};
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    uint32_t v = vals[q];
#pragma unroll
    for(int d = 16; d; d >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, d);
    if((tid & 31) == 0 && v)
      atomicAdd(&p.stats[q == 0 ? STAT_FRAGMENTS : (q == 1 ? STAT_STORED : (q == 2 ? STAT_TAIL : STAT_OPAQUE))], (unsigned long long)v);
  }
}

// ---- dispatch ----------------------------------------------------------------------------------------------------------
template <int PASS>
static void launchPass(const FrameParams& p, cudaStream_t s)
{
  const unsigned grid = (unsigned)(p.tilesX * p.tileRowsLocal);
  const bool     ss   = p.sampleShading != 0;
  if(p.msaa == 1)
    k_raster<PASS, 1, false><<<grid, RASTER_THREADS, 0, s>>>(p);
  else if(p.msaa == 4)
  {
    if(ss)
      k_raster<PASS, 4, true><<<grid, RASTER_THREADS, 0, s>>>(p);
    else
      k_raster<PASS, 4, false><<<grid, RASTER_THREADS, 0, s>>>(p);
  }
  else
  {
    if(ss)
      k_raster<PASS, 8, true><<<grid, RASTER_THREADS, 0, s>>>(p);
    else
      k_raster<PASS, 8, false><<<grid, RASTER_THREADS, 0, s>>>(p);
  }
}

int launchRaster(const FrameParams& p, int pass, cudaStream_t s)
{
  if(p.tilesX * p.tileRowsLocal == 0)
    return 0;
  switch(pass)
  {
    case PASS_SIMPLE: launchPass<PASS_SIMPLE>(p, s); break;
    case PASS_LINKEDLIST: launchPass<PASS_LINKEDLIST>(p, s); break;
    case PASS_LOOP_COLOR: launchPass<PASS_LOOP_COLOR>(p, s); break;
    case PASS_LOOP64: launchPass<PASS_LOOP64>(p, s); break;
    case PASS_SPINLOCK: launchPass<PASS_SPINLOCK>(p, s); break;
    case PASS_INTERLOCK: launchPass<PASS_INTERLOCK>(p, s); break;
    case PASS_WEIGHTED: launchPass<PASS_WEIGHTED>(p, s); break;
    case PASS_LOOP_DEPTH: launchPass<PASS_LOOP_DEPTH>(p, s); break;
    case PASS_OPAQUE: launchPass<PASS_OPAQUE>(p, s); break;
    default: return 0;
  }
  return 1;
}

}  // namespace oit
