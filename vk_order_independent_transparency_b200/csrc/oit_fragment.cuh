// oit_fragment.cuh -- triangle slot, varyings/shading and the per-technique fragment ("colour pass") programs.
// Included by oit_raster.cu only.  Reference citations are on each function.
#pragma once
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
  float    rarea;         // 1 / (float)area2
  uint32_t box;           // bx0 | by0 << 4 | (bw-1) << 8 | (bh-1) << 12 | bias bits << 16 | zSafe << 19 | small << 20
  uint32_t rcpW;          // ceil(65536 / bw)
};

struct FragCtx
{
  const FrameParams& p;
  const SrgbTables&  t;
  uint32_t           nFrag, nStored, nTail, nOpaque;
  // WBOIT targets of the tile in shared memory (fused frame kernel) or nullptr = the global RGBA16F / R16F images
  uint2*    wAccTile;
  uint16_t* wRevTile;
  // the A-buffer slice + aux words the fragment programs work on: the global buffers (viewSize = W * localH, pixel index
  // = yl * W + x), or -- fused frame kernel, k-buffer techniques -- the tile's slice in shared memory (viewSize = 256,
  // pixel index = position inside the tile), addressed with the reference's own index arithmetic either way
  uint32_t* abuf;
  uint32_t* aux;
  uint32_t* adepth;
  uint32_t* spin;
  size_t    viewSize;
  bool      onChip;
};

// loads of words that atomics of this kernel modify: L2 (ld.cg) for the global buffers, plain for the shared-memory slice
__device__ __forceinline__ uint32_t ldcg32(const FragCtx& c, const uint32_t* a) { return c.onChip ? *a : __ldcg(a); }
__device__ __forceinline__ unsigned long long ldcg64(const FragCtx& c, const unsigned long long* a) { return c.onChip ? *a : __ldcg(a); }

// ---- varyings + shading -------------------------------------------------------------------------------------------
struct Bary
{
  float l0, l1, l2;
};
// screen-space barycentrics from the (unbiased) edge functions 1 and 2, already converted to float
__device__ __forceinline__ Bary makeBary(float e1, float e2, float rarea)
{
  Bary b;
  b.l1 = __fmul_rn(e1, rarea);
  b.l2 = __fmul_rn(e2, rarea);
  b.l0 = __fsub_rn(__fsub_rn(1.0f, b.l1), b.l2);
  return b;
}
__device__ __forceinline__ float depthAt(const TriSlot& s, const Bary& b)
{
  return clamp01(__fmaf_rn(b.l2, s.dz2, __fmaf_rn(b.l1, s.dz1, s.z0)));
}

// TriSlot::box flags of a piece of a near-clipped triangle: vidx[0] is the index of its ClipEntry (the frame's clip table,
// oit_clip.cuh / k_bin_emit), the geometry (x, y, z, iw) is the piece's
constexpr uint32_t SLOT_CLIPPED = 1u << 21;  // the slot is a piece of a near-clipped triangle
constexpr uint32_t SLOT_SWAPPED = 1u << 23;  // its vertices 1 and 2 were exchanged to make area2 > 0

// shading() + goochLighting() (shaderCommon.glsl:36-56) of the interpolated normal (v[0..2]) and colour (v[3..6])
__device__ __forceinline__ Color4 shadeVaryings(const DeviceUbo* ubo, const float v[7])
{
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
  c.a = clamp01(__fmaf_rn(v[6], ubo->alphaWidth, ubo->alphaMin));
  return c;
}

// Interpolants (shaderCommon.glsl:25-31) perspective-correct at `b`, then shading() (shaderCommon.glsl:36-56)
template <bool NEED_VIEWZ>
__device__ __forceinline__ Color4 shadeAt(const FrameParams& p, const TriSlot& s, const Bary& b, float& viewz)
{
  const float q0 = __fmul_rn(b.l0, s.iw[0]), q1 = __fmul_rn(b.l1, s.iw[1]), q2 = __fmul_rn(b.l2, s.iw[2]);
  const float rden = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(q0, q1), q2));
  // a piece of a near-clipped triangle reads its three vertex records (and view depths) from its ClipEntry; slot vertex
  // k is piece vertex k, or 3 - k for k > 0 when the set-up exchanged vertices 1 and 2
  const bool       clipped = (s.box & SLOT_CLIPPED) != 0;
  const ClipEntry* ce      = p.clipEntries + s.vidx[0];
  const int        m1      = (s.box & SLOT_SWAPPED) ? 2 : 1;
  const float4*    a0      = clipped ? reinterpret_cast<const float4*>(ce->attr[0]) : p.tvAttr + 2 * (size_t)s.vidx[0];
  const float4*    a1      = clipped ? reinterpret_cast<const float4*>(ce->attr[m1]) : p.tvAttr + 2 * (size_t)s.vidx[1];
  const float4*    a2      = clipped ? reinterpret_cast<const float4*>(ce->attr[3 - m1]) : p.tvAttr + 2 * (size_t)s.vidx[2];
  // normal + colour + view depth of each vertex: two 128-bit loads (written by the vertex stage / the binning of this frame)
  float v0[8], v1[8], v2[8];
  auto  fetch = [](const float4* a, float* o) {
    const float4 x = __ldg(a), y = __ldg(a + 1);
    o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w; o[4] = y.x; o[5] = y.y; o[6] = y.z; o[7] = y.w;
  };
  fetch(a0, v0);
  fetch(a1, v1);
  fetch(a2, v2);
  const float vz0 = v0[7], vz1 = v1[7], vz2 = v2[7];
  float v[7];
#pragma unroll
  for(int k = 0; k < 7; k++)
    v[k] = __fmul_rn(__fmaf_rn(q2, v2[k], __fmaf_rn(q1, v1[k], __fmul_rn(q0, v0[k]))), rden);
  if(NEED_VIEWZ)
    viewz = __fmul_rn(__fmaf_rn(q2, vz2, __fmaf_rn(q1, vz1, __fmul_rn(q0, vz0))), rden);
  return shadeVaryings(p.ubo, v);
}

// ---- fragment programs ----------------------------------------------------------------------------------------------
// All return the colour handed to the ROP (premultiplied; zero = no-op).  x, yl: pixel (yl = row inside this band's
// buffers); pix = yl * W + x; ai = aux index of (sampleID, pixel).

// K2 oitSimple.frag.glsl:50-91
// `old` = imageAtomicAdd(imgAux, coord, 1u), issued by preInvoke() before the shading so that its latency overlaps
__device__ __forceinline__ Color4 fragSimple(FragCtx& c, size_t pix, uint32_t old, uint32_t sampleID, uint32_t mask, const Color4& rgba, float z)
{
  const FrameParams& p        = c.p;
  const size_t       viewSize = c.viewSize;
  const size_t       listPos  = viewSize * p.L * sampleID + pix;
  if(old < (uint32_t)p.L)
  {
    const uint32_t packed = packColor(c.t, rgba);
    if(p.coverage)
      reinterpret_cast<uint4*>(c.abuf)[listPos + (size_t)old * viewSize] = make_uint4(packed, __float_as_uint(z), mask, 0u);
    else
      reinterpret_cast<uint2*>(c.abuf)[listPos + (size_t)old * viewSize] = make_uint2(packed, __float_as_uint(z));
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

// the linked-list allocator (oitLinkedList.frag.glsl:55): the single-address counter is bumped once per converged
// warp group (warp-aggregated atomicAdd); called before the shading so that the L2 round trip overlaps with it
__device__ __forceinline__ uint32_t allocLinkedListNode(const FrameParams& p)
{
  cg::coalesced_group g    = cg::coalesced_threads();
  uint32_t            base = 0;
  if(g.thread_rank() == 0)
    base = atomicAdd(p.counter, g.size());
  return g.shfl(base, 0) + g.thread_rank() + 1u;
}

// K4 oitLinkedList.frag.glsl:51-85.  The head exchange runs with the pixel exclusively owned (tile-ordered arbitration),
// so imageAtomicExchange is a plain load + store that stays in this SM's L1.
__device__ __forceinline__ Color4 fragLinkedList(FragCtx& c, size_t ai, uint32_t newOffset, uint32_t mask, const Color4& rgba, float z)
{
  const FrameParams& p = c.p;
  if(newOffset >= p.capacity)
  {
    if(p.tailBlend)
    {
      c.nTail++;
      return premultiply(rgba);
    }
    return zeroColor();
  }
  const uint32_t oldOffset = c.aux[ai];
  c.aux[ai]                = newOffset;
  reinterpret_cast<uint4*>(c.abuf)[newOffset] =
      make_uint4(packColor(c.t, rgba), __float_as_uint(z), p.coverage ? mask : 0u, oldOffset);
  c.nStored++;
  return zeroColor();
}

// K6 oitLoop.frag.glsl:57-100 (depth pass)
__device__ __forceinline__ void fragLoopDepth(FragCtx& c, size_t pix, uint32_t sampleID, float z)
{
  const FrameParams& p        = c.p;
  const size_t       viewSize = c.viewSize;
  uint32_t*          list     = c.abuf + viewSize * p.L * 2 * sampleID + pix;
  uint32_t           zcur     = __float_as_uint(z);
  int                i        = 0;
  uint32_t           pretest  = ldcg32(c, list + (size_t)(p.L - 1) * viewSize);
  if(zcur > pretest)
    return;
  pretest = ldcg32(c, list + (size_t)(p.L / 2) * viewSize);
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
  const size_t       viewSize = c.viewSize;
  uint32_t*          list     = c.abuf + viewSize * p.L * 2 * sampleID + pix;
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
  const size_t        viewSize = c.viewSize;
  unsigned long long* list     = reinterpret_cast<unsigned long long*>(c.abuf) + viewSize * p.L * sampleID + pix;
  unsigned long long  zcur     = ((unsigned long long)__float_as_uint(z) << 32) | packColor(c.t, rgba);
  int                 i        = 0;
  bool                canInsert = true;
  unsigned long long  pretest   = ldcg64(c, list + (size_t)(p.L - 1) * viewSize);
  if(zcur > pretest)
    canInsert = false;
  else
  {
    pretest = ldcg64(c, list + (size_t)(p.L / 2) * viewSize);
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
  const size_t       viewSize = c.viewSize;
  const size_t       listPos  = viewSize * p.L * sampleID + pix;
  const uint32_t     oldCounter = c.aux[ai];
  c.aux[ai]                     = oldCounter + 1u;
  if(oldCounter < (uint32_t)p.L)
  {
    if(p.coverage)
      reinterpret_cast<uint4*>(c.abuf)[listPos + (size_t)oldCounter * viewSize] = make_uint4(packed, zbits, mask, 0u);
    else
      reinterpret_cast<uint2*>(c.abuf)[listPos + (size_t)oldCounter * viewSize] = make_uint2(packed, zbits);
    color = zeroColor();
    return true;
  }
  int      furthest = 0;
  uint32_t maxDepth = 0;
  for(int i = 0; i < p.L; i++)
  {
    const size_t   e         = listPos + (size_t)i * viewSize;
    const uint32_t testDepth = p.coverage ? c.abuf[e * 4 + 1] : c.abuf[e * 2 + 1];
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
      color                               = unpackColor(c.t, c.abuf[e * 4]);
      reinterpret_cast<uint4*>(c.abuf)[e] = make_uint4(packed, zbits, mask, 0u);
    }
    else
    {
      color                               = unpackColor(c.t, c.abuf[e * 2]);
      reinterpret_cast<uint2*>(c.abuf)[e] = make_uint2(packed, zbits);
    }
    c.adepth[ai] = maxDepth;
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
    const uint32_t oldDepth = ldcg32(c, &c.adepth[ai]);  // racy-but-conservative early-out outside the lock (:67-68)
    if(zbits <= oldDepth)
    {
      bool done = mask == 0;
      while(!done)
      {
        const uint32_t old = atomicExch(&c.spin[ai], 1u);
        if(old == 0u)
        {
          stored = lockCriticalSection(c, pix, ai, sampleID, mask, packed, zbits, color);
          __threadfence();
          atomicExch(&c.spin[ai], 0u);
          done = true;
        }
      }
    }
  }
  else
  {
    // beginInvocationInterlock .. endInvocationInterlock: the arbitration loop of the caller IS the ordered interlock
    if(zbits <= c.adepth[ai])
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

// K15 oitWeighted.frag.glsl:53-79 + BlendMode::WEIGHTED_COLOR into RGBA16F / R16F (main.cpp:559-575), in two halves:
// what the shader outputs (the weighted premultiplied colour, and 1 - alpha for the revealage target) ...
__device__ __forceinline__ void weightedSource(const Color4& rgba, float viewz, float src[4], float& om)
{
  const Color4 col        = premultiply(rgba);
  const float  depthZ     = __fmul_rn(-viewz, 10.0f);
  const float  x          = __fdiv_rn(depthZ, 200.0f);
  const float  x2         = __fmul_rn(x, x);
  const float  x4         = __fmul_rn(x2, x2);
  float        distWeight = __fdiv_rn(0.03f, __fadd_rn(1e-5f, x4));
  distWeight              = distWeight < 1e-2f ? 1e-2f : (distWeight > 3e3f ? 3e3f : distWeight);
  const float mx     = fmaxf(fmaxf(col.r, col.g), fmaxf(col.b, col.a));
  float       aw     = fminf(1.0f, __fmaf_rn(mx, 40.0f, 0.01f));
  aw                 = __fmul_rn(aw, aw);
  const float weight = __fmul_rn(aw, distWeight);
  om                 = __fsub_rn(1.0f, col.a);
  src[0]             = __fmul_rn(col.r, weight);
  src[1]             = __fmul_rn(col.g, weight);
  src[2]             = __fmul_rn(col.b, weight);
  src[3]             = __fmul_rn(col.a, weight);
}
// ... and the ROP on one covered sample: add on RGBA16F (fp16 -> fp32, add, round to nearest even back to fp16, two channels
// per conversion), multiply on R16F
__device__ __forceinline__ void weightedBlendSample(uint2* acc, uint16_t* rev, const float src[4], float om)
{
  const uint2   a   = *acc;
  const float2  rg  = __half22float2(*reinterpret_cast<const __half2*>(&a.x));
  const float2  ba  = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
  const __half2 nrg = __floats2half2_rn(__fadd_rn(rg.x, src[0]), __fadd_rn(rg.y, src[1]));
  const __half2 nba = __floats2half2_rn(__fadd_rn(ba.x, src[2]), __fadd_rn(ba.y, src[3]));
  *acc              = make_uint2(*reinterpret_cast<const uint32_t*>(&nrg), *reinterpret_cast<const uint32_t*>(&nba));
  *rev              = f2h(__fmul_rn(h2f(*rev), om));
}
template <int S>
__device__ __forceinline__ void fragWeighted(FragCtx& c, size_t pix, int pl, uint32_t mask, const Color4& rgba, float viewz)
{
  const FrameParams& p = c.p;
  float              src[4], om;
  weightedSource(rgba, viewz, src, om);
  // the pixel's S accumulator / revealage samples: shared-memory tile (fused frame) or the global images
  uint2*    acc = c.wAccTile ? c.wAccTile + pl * S : reinterpret_cast<uint2*>(p.wacc) + pix * S;
  uint16_t* rev = c.wRevTile ? c.wRevTile + pl * S : p.wrev + pix * S;
#pragma unroll
  for(int s = 0; s < S; s++)
    if(mask & (1u << s))
      weightedBlendSample(acc + s, rev + s, src, om);
  c.nStored++;
}

// Code that most frames never execute (the tail-blend ROP, 64-bit coverage of huge triangles, the per-sample depth test
// against opaque geometry) is kept out of line: the frame kernel is instruction-cache bound (ncu: "no instruction" stalls),
// and inlined S-times-unrolled copies of these paths sat in the middle of its hot loops.
#ifndef OIT_COLD_NOINLINE
#define OIT_COLD_NOINLINE 1
#endif
#if OIT_COLD_NOINLINE
#define OIT_COLD __noinline__
#else
#define OIT_COLD __forceinline__
#endif

// px: the S colour samples of the pixel -- in m_colorImage, or in the shared-memory tile of the fused frame kernel
template <int S>
__device__ OIT_COLD void ropSamplesNonZero(const SrgbTables& t, uint32_t* px, uint32_t mask, Color4 src)
{
#pragma unroll 1
  for(int s = 0; s < S; s++)
    if(mask & (1u << s))
      px[s] = ropPremult(t, px[s], src);
}
template <int S>
__device__ __forceinline__ void ropSamples(const FragCtx& c, uint32_t* px, uint32_t mask, const Color4& src)
{
  if(isZero(src))
    return;  // identity blend: encode(decode(v)) == v for every 8-bit v
  ropSamplesNonZero<S>(c.t, px, mask, src);
}

// the part of an invocation that does not depend on the shaded colour: issued first so that its memory latency is
// hidden behind the interpolation + shading arithmetic
template <int PASS>
__device__ __forceinline__ uint32_t preInvoke(const FragCtx& c, size_t pixA, uint32_t sampleID)
{
  if(PASS == PASS_LINKEDLIST)
    return allocLinkedListNode(c.p);
  if(PASS == PASS_SIMPLE)
    return atomicAdd(&c.aux[(size_t)sampleID * c.viewSize + pixA], 1u);
  return 0u;
}

// one colour-pass invocation + its ROP write
template <int PASS, int S>
__device__ __forceinline__ void invoke(FragCtx& c, int x, int yl, uint32_t sampleID, uint32_t mask, const Color4& rgba, float z, float viewz,
                                       uint32_t token, uint32_t* colorPx, int pl)
{
  const FrameParams& p    = c.p;
  const size_t       pixG = (size_t)yl * p.W + x;            // pixel index in the global images
  const size_t       pix  = c.onChip ? (size_t)pl : pixG;    // pixel index in the A-buffer slice being used
  const size_t       ai   = (size_t)sampleID * c.viewSize + pix;
  if(PASS == PASS_LOOP_DEPTH)
  {
    fragLoopDepth(c, pix, sampleID, z);
    return;
  }
  c.nFrag++;
  Color4 out = zeroColor();
  switch(PASS)
  {
    case PASS_SIMPLE: out = fragSimple(c, pix, token, sampleID, mask, rgba, z); break;
    case PASS_LINKEDLIST: out = fragLinkedList(c, ai, token, mask, rgba, z); break;
    case PASS_LOOP_COLOR: out = fragLoopColor(c, pix, sampleID, rgba, z); break;
    case PASS_LOOP64: out = fragLoop64(c, pix, sampleID, rgba, z); break;
    case PASS_SPINLOCK: out = fragLock<true>(c, pix, ai, sampleID, mask, rgba, z); break;
    case PASS_INTERLOCK: out = fragLock<false>(c, pix, ai, sampleID, mask, rgba, z); break;
    case PASS_WEIGHTED: fragWeighted<S>(c, pixG, pl, mask, rgba, viewz); return;
  }
  ropSamples<S>(c, colorPx, mask, out);
}

// ---- the same programs with the pixel EXCLUSIVELY OWNED and its fragments arriving in primitive order -----------------------
// (oit_raster_q.cu: one owner thread per pixel and batch walks the pixel's fragments in order.)  Mutual exclusion is
// structural there, so the atomics of the GLSL become plain read-modify-writes on the owner's slice; what each program leaves
// in the A-buffer and hands to the ROP is what the cascade of atomics leaves under the sequential schedule.

// K6 (oitLoop.frag.glsl:57-100): the L smallest DISTINCT depths, ascending.  The atomicMin cascade = a sorted insert that
// carries the displaced value on, stops at an empty slot or at an equal depth, and drops what falls off the end.
__device__ __forceinline__ void ownedLoopDepth(FragCtx& c, size_t pix, uint32_t sampleID, float z)
{
  const FrameParams& p        = c.p;
  const size_t       viewSize = c.viewSize;
  uint32_t*          list     = c.abuf + viewSize * p.L * 2 * sampleID + pix;
  uint32_t           zcur     = __float_as_uint(z);
  if(zcur > list[(size_t)(p.L - 1) * viewSize])
    return;
  for(int i = 0; i < p.L; i++)
  {
    const uint32_t ztest = list[(size_t)i * viewSize];
    if(ztest == zcur)
      return;
    if(ztest > zcur)
    {
      list[(size_t)i * viewSize] = zcur;
      if(ztest == 0xFFFFFFFFu)
        return;
      zcur = ztest;
    }
  }
}

// K9 (oitLoop64.frag.glsl:65-141): the L smallest depth|colour keys, ascending; a key that does not fit -- the new one, or
// the largest one it displaces -- is tail blended
__device__ __forceinline__ Color4 ownedLoop64(FragCtx& c, size_t pix, uint32_t sampleID, const Color4& rgba, float z)
{
  const FrameParams&  p        = c.p;
  const size_t        viewSize = c.viewSize;
  unsigned long long* list     = reinterpret_cast<unsigned long long*>(c.abuf) + viewSize * p.L * sampleID + pix;
  unsigned long long  zcur     = ((unsigned long long)__float_as_uint(z) << 32) | packColor(c.t, rgba);
  if(!(zcur > list[(size_t)(p.L - 1) * viewSize]))
    for(int i = 0; i < p.L; i++)
    {
      const unsigned long long ztest = list[(size_t)i * viewSize];
      if(ztest > zcur)
      {
        list[(size_t)i * viewSize] = zcur;
        if(ztest == ~0ull)
        {
          c.nStored++;
          return zeroColor();
        }
        zcur = ztest;
      }
    }
  if(p.tailBlend)
  {
    c.nTail++;
    return premultiply(unpackColor(c.t, (uint32_t)(zcur & 0xFFFFFFFFull)));
  }
  return zeroColor();
}

// one colour-pass invocation of an owned pixel; returns the colour it hands to the ROP (premultiplied; zero = nothing)
template <int PASS, int S>
__device__ __forceinline__ Color4 invokeOwned(FragCtx& c, int x, int yl, uint32_t sampleID, uint32_t mask, const Color4& rgba, float z, float viewz,
                                              int pl)
{
  const FrameParams& p    = c.p;
  const size_t       pixG = (size_t)yl * p.W + x;          // pixel index in the global images
  const size_t       pix  = c.onChip ? (size_t)pl : pixG;  // pixel index in the A-buffer slice being used
  const size_t       ai   = (size_t)sampleID * c.viewSize + pix;
  if(PASS == PASS_LOOP_DEPTH)
  {
    ownedLoopDepth(c, pix, sampleID, z);
    return zeroColor();
  }
  c.nFrag++;
  Color4 out = zeroColor();
  switch(PASS)
  {
    case PASS_SIMPLE: {
      const uint32_t old = c.aux[ai];  // imageAtomicAdd(imgAux, coord, 1u)
      c.aux[ai]          = old + 1u;
      out                = fragSimple(c, pix, old, sampleID, mask, rgba, z);
      break;
    }
    case PASS_LINKEDLIST: out = fragLinkedList(c, ai, allocLinkedListNode(p), mask, rgba, z); break;
    case PASS_LOOP_COLOR: out = fragLoopColor(c, pix, sampleID, rgba, z); break;
    case PASS_LOOP64: out = ownedLoop64(c, pix, sampleID, rgba, z); break;
    case PASS_SPINLOCK:   // the lock word is never contended: the critical section runs as under the ordered interlock
    case PASS_INTERLOCK: out = fragLock<false>(c, pix, ai, sampleID, mask, rgba, z); break;
    case PASS_WEIGHTED: fragWeighted<S>(c, pixG, pl, mask, rgba, viewz); break;
  }
  return out;
}

}  // namespace oit
