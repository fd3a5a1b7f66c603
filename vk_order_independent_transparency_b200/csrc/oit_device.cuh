// oit_device.cuh -- device-side arithmetic shared by the raster, composite and resolve kernels.
//
// Arithmetic contract (DESIGN.md "arith spec"): every float expression is spelled with explicit round-to-nearest
// intrinsics (__fmaf_rn, __fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn) and the library is compiled with -fmad=false,
// so results do not depend on the compiler's contraction choices.  The sRGB conversions of the reference GLSL
// (shaderCommon.glsl:60-104) only ever see 8-bit codes, so they are exact table operations here.
#pragma once
#include <cuda_fp16.h>

#include "oit_internal.h"

namespace oit {

struct __align__(16) SrgbTables
{
  float   dec[256];  // sRGB8 code -> linear
  float   thr[260];  // thr[k]: smallest linear value whose code is >= k; thr[0] = -inf, thr[256..259] = +inf
  float   a255[256]; // v / 255.0f
  uint8_t bucket[SRGB_BUCKET_BYTES];
};
static_assert(sizeof(SrgbTables) == SRGB_TABLE_BYTES && SRGB_TABLE_BYTES % 16 == 0, "layout shared with buildTables (oit_api.cu)");

// (sm must be 16-byte aligned; g is the start of a cudaMalloc allocation)
__device__ __forceinline__ void loadTables(SrgbTables& sm, const float* __restrict__ g)
{
  for(int i = threadIdx.x; i < (int)(SRGB_TABLE_BYTES / 16); i += blockDim.x)
    reinterpret_cast<float4*>(&sm)[i] = __ldg(reinterpret_cast<const float4*>(g) + i);
}

__device__ __forceinline__ float clamp01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }

// 8-bit sRGB code of a linear value = the largest k with c >= thr[k]: the code of the value's bucket, plus one if the value
// has passed the one threshold that can lie inside the bucket.  Exact by construction (no transcendental guess).
__device__ __forceinline__ uint32_t enc8(const SrgbTables& t, float c)
{
  const float    cc  = fminf(fmaxf(c, 0.f), 1.f);  // NaN -> 0
  const uint32_t idx = max(__float_as_uint(cc) >> 16, SRGB_BUCKET_BASE) - SRGB_BUCKET_BASE;
  const uint32_t k   = t.bucket[idx];
  return k + (cc >= t.thr[k + 1] ? 1u : 0u);
}
__device__ __forceinline__ uint32_t unorm8(float a) { return __float2uint_rn(__fmul_rn(clamp01(a), 255.0f)); }

struct Color4
{
  float r, g, b, a;
};
__device__ __forceinline__ Color4 zeroColor() { return Color4{0.f, 0.f, 0.f, 0.f}; }
__device__ __forceinline__ bool   isZero(const Color4& c) { return c.r == 0.f && c.g == 0.f && c.b == 0.f && c.a == 0.f; }

// packUnorm4x8(unPremultLinearToSRGB(c)): r in bits 0-7, a in bits 24-31 (oitSimple.frag.glsl:55,70)
__device__ __forceinline__ uint32_t packColor(const SrgbTables& t, const Color4& c)
{
  return enc8(t, c.r) | (enc8(t, c.g) << 8) | (enc8(t, c.b) << 16) | (unorm8(c.a) << 24);
}
// unPremultSRGBToLinear(unpackUnorm4x8(p)) (shaderCommon.glsl:84-104)
__device__ __forceinline__ Color4 unpackColor(const SrgbTables& t, uint32_t p)
{
  return Color4{t.dec[p & 255], t.dec[(p >> 8) & 255], t.dec[(p >> 16) & 255], t.a255[p >> 24]};
}
__device__ __forceinline__ Color4 premultiply(const Color4& c)
{
  return Color4{__fmul_rn(c.r, c.a), __fmul_rn(c.g, c.a), __fmul_rn(c.b, c.a), c.a};
}
// doBlend (shaderCommon.glsl:108-112): color over base, both premultiplied
__device__ __forceinline__ void doBlend(Color4& color, const Color4& base)
{
  const float t = __fsub_rn(1.0f, color.a);
  color.r       = __fmaf_rn(t, base.r, color.r);
  color.g       = __fmaf_rn(t, base.g, color.g);
  color.b       = __fmaf_rn(t, base.b, color.b);
  color.a       = __fmaf_rn(t, base.a, color.a);
}
// doBlendPacked (shaderCommon.glsl:117-124)
__device__ __forceinline__ void doBlendPacked(const SrgbTables& t, Color4& color, uint32_t packed)
{
  doBlend(color, premultiply(unpackColor(t, packed)));
}

// ---- ROP on the B8G8R8A8_SRGB colour target (oit.cpp:58; blend states main.cpp:540-592) ----------------------------
__device__ __forceinline__ Color4 decodeDst(const SrgbTables& t, uint32_t d)
{
  return Color4{t.dec[(d >> 16) & 255], t.dec[(d >> 8) & 255], t.dec[d & 255], t.a255[d >> 24]};
}
// One copy per kernel instead of one per call site (OIT_SHARE_ENCODE): three enc8 are ~75 SASS instructions, the frame
// kernel has a dozen call sites (ROP, composite, resolve) and is instruction-cache bound; the call costs less than the misses.
#ifndef OIT_SHARE_ENCODE
#define OIT_SHARE_ENCODE 1
#endif
#if OIT_SHARE_ENCODE
static __device__ __noinline__ uint32_t encodeDst(const SrgbTables& t, Color4 c)
#else
__device__ __forceinline__ uint32_t encodeDst(const SrgbTables& t, const Color4& c)
#endif
{
  return enc8(t, c.b) | (enc8(t, c.g) << 8) | (enc8(t, c.r) << 16) | (unorm8(c.a) << 24);
}
// BlendMode::PREMULTIPLIED: dst = src + (1 - src.a) * dst (main.cpp:548-558)
__device__ __forceinline__ uint32_t ropPremult(const SrgbTables& t, uint32_t dst, const Color4& src)
{
  const Color4 d  = decodeDst(t, dst);
  const float  om = __fsub_rn(1.0f, src.a);
  return encodeDst(t, Color4{__fmaf_rn(om, d.r, src.r), __fmaf_rn(om, d.g, src.g), __fmaf_rn(om, d.b, src.b),
                             __fmaf_rn(om, d.a, src.a)});
}
// BlendMode::WEIGHTED_COMPOSITE: (1 - src.a) * src + src.a * dst (main.cpp:576-588)
__device__ __forceinline__ uint32_t ropWeightedComposite(const SrgbTables& t, uint32_t dst, const Color4& src)
{
  const Color4 d  = decodeDst(t, dst);
  const float  om = __fsub_rn(1.0f, src.a);
  return encodeDst(t, Color4{__fmaf_rn(src.a, d.r, __fmul_rn(om, src.r)), __fmaf_rn(src.a, d.g, __fmul_rn(om, src.g)),
                             __fmaf_rn(src.a, d.b, __fmul_rn(om, src.b)), __fmaf_rn(src.a, d.a, __fmul_rn(om, src.a))});
}

__device__ __forceinline__ float    h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

// standard sample locations in 1/256 px (Vulkan standardSampleLocations)
__device__ __forceinline__ void samplePos(int S, int s, int& sx, int& sy)
{
  if(S == 1)
  {
    sx = 128;
    sy = 128;
  }
  else if(S == 4)
  {
    const int X[4] = {96, 224, 32, 160}, Y[4] = {32, 96, 160, 224};
    sx = X[s];
    sy = Y[s];
  }
  else
  {
    const int X[8] = {144, 112, 208, 80, 48, 16, 176, 240}, Y[8] = {80, 176, 144, 48, 208, 112, 240, 16};
    sx = X[s];
    sy = Y[s];
  }
}

// ---- CTA-wide exclusive scan (warp shuffles + one smem hop) -------------------------------------------------------------
__device__ __forceinline__ uint32_t warpInclusiveScan(uint32_t v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for(int d = 1; d < 32; d <<= 1)
  {
    const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
    if(lane >= d)
      v += n;
  }
  return v;
}
// exclusive scan across the CTA; returns the exclusive prefix of `v`, total in `total`. smem: >= 33 words
__device__ __forceinline__ uint32_t blockExclusiveScan(uint32_t v, uint32_t* smem, uint32_t& total)
{
  const int      lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t inc = warpInclusiveScan(v);
  if(lane == 31)
    smem[warp] = inc;
  __syncthreads();
  if(warp == 0)
  {
    uint32_t w  = lane < nw ? smem[lane] : 0;
    uint32_t wi = warpInclusiveScan(w);
    smem[lane]  = wi - w;
    if(lane == 31)
      smem[32] = wi;
  }
  __syncthreads();
  const uint32_t r = smem[warp] + inc - v;
  total            = smem[32];
  __syncthreads();
  return r;
}

}  // namespace oit
