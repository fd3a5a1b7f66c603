// oit_fused.cuh -- composite + resolve of ONE tile by the CTA that has just finished the tile's colour pass.
//
// Same programs as oit_composite.cu (K3/K5/K8/K10/K12/K14/K16 + copyOffscreenToBackBuffer, cited there), written for
// OIT_LAYERS <= 8 with the per-pixel fragment arrays in shared memory (element i of thread t at [i][t]: conflict free).
// Because the tile's list nodes / k-buffer slots were written by this SM moments ago they are read back from L1/L2, and
// because the tile's colour samples live in shared memory the composite's ROP read-modify-write and the resolve never
// touch HBM: only the resolved BGRA8 pixel is written.
#pragma once
#include "oit_device.cuh"

namespace oit {

constexpr int FUSED_LCAP = 8;  // the fused frame kernel is used for OIT_LAYERS <= 8; larger values take the staged kernels

#ifndef OIT_COMPACT_COMPOSITE
#define OIT_COMPACT_COMPOSITE 1
#endif

struct FusedArrays
{
  uint32_t c[FUSED_LCAP][TILE_PIX];
  uint32_t d[FUSED_LCAP][TILE_PIX];
  uint32_t m[FUSED_LCAP][TILE_PIX];
};

__device__ __forceinline__ bool fusedGE(uint32_t a, uint32_t b) { return __uint_as_float(a) >= __uint_as_float(b); }
__device__ __forceinline__ bool fusedLT(uint32_t a, uint32_t b) { return __uint_as_float(a) < __uint_as_float(b); }

// bubbleSort of oitCompositeDefines.glsl:51-89 on the first n entries of thread t's column (swap on >=)
__device__ __forceinline__ void fusedBubbleSort(FusedArrays& A, int t, int n)
{
  for(int i = n - 2; i >= 0; --i)
    for(int j = 0; j <= i; ++j)
      if(fusedGE(A.d[j][t], A.d[j + 1][t]))
      {
        const uint32_t c = A.c[j + 1][t], d = A.d[j + 1][t], m = A.m[j + 1][t];
        A.c[j + 1][t] = A.c[j][t];
        A.d[j + 1][t] = A.d[j][t];
        A.m[j + 1][t] = A.m[j][t];
        A.c[j][t]     = c;
        A.d[j][t]     = d;
        A.m[j][t]     = m;
      }
}

// insertionSortTail / insertionSort (oitCompositeDefines.glsl:94-139); returns the colour that falls out
template <bool TAIL>
__device__ __forceinline__ uint32_t fusedInsert(FusedArrays& A, int t, int L, uint32_t c, uint32_t d, uint32_t m)
{
  uint32_t outColor = c;
  if(!TAIL || fusedLT(d, A.d[L - 1][t]))
  {
    for(int i = 0; i < L; ++i)
      if(fusedLT(d, A.d[i][t]))
      {
        outColor = A.c[L - 1][t];
        for(int j = L - 1; j > i; j--)
        {
          A.c[j][t] = A.c[j - 1][t];
          A.d[j][t] = A.d[j - 1][t];
          A.m[j][t] = A.m[j - 1][t];
        }
        A.c[i][t] = c;
        A.d[i][t] = d;
        A.m[i][t] = m;
        break;
      }
  }
  return outColor;
}

// blend of the sorted fragments (oitSimple.frag.glsl:138-167).  Coverage mode keeps one accumulator per sample and walks
// the fragments once: every sample still sees its fragments front to back, so the arithmetic is the reference's.
template <int S>
__device__ __forceinline__ Color4 fusedBlend(const SrgbTables& tb, const FusedArrays& A, int t, int n, bool coverage)
{
  if(coverage && S > 1)
  {
    Color4 sc[S];
#pragma unroll
    for(int s = 0; s < S; s++)
      sc[s] = zeroColor();
#if OIT_COMPACT_COMPOSITE
#pragma unroll 1  // the S-wide body is big enough: unrolling over the fragments only costs instruction-cache space
#endif
    for(int i = 0; i < n; i++)
    {
      const Color4   pm = premultiply(unpackColor(tb, A.c[i][t]));
      const uint32_t m  = A.m[i][t];
#pragma unroll
      for(int s = 0; s < S; s++)
        if(m & (1u << s))
          doBlend(sc[s], pm);
    }
    Color4 sum = zeroColor();
#pragma unroll
    for(int s = 0; s < S; s++)
    {
      sum.r = __fadd_rn(sum.r, sc[s].r);
      sum.g = __fadd_rn(sum.g, sc[s].g);
      sum.b = __fadd_rn(sum.b, sc[s].b);
      sum.a = __fadd_rn(sum.a, sc[s].a);
    }
    const float inv = 1.0f / (float)S;
    return Color4{__fmul_rn(sum.r, inv), __fmul_rn(sum.g, inv), __fmul_rn(sum.b, inv), __fmul_rn(sum.a, inv)};
  }
  Color4 sum = zeroColor();
  for(int i = 0; i < n; i++)
    doBlendPacked(tb, sum, A.c[i][t]);
  return sum;
}

// the composite invocation of one (pixel, sampleID): returns the colour handed to the ROP
// the A-buffer slice the composite reads: the global buffers, or the tile's shared-memory slice of the k-buffer techniques
struct AbufView
{
  const uint32_t* abuf;
  const uint32_t* aux;
  size_t          viewSize;
};

// pix: the pixel's index inside that slice
// ALG: the technique, a compile-time constant of the kernel instance (the other techniques' code is not instantiated)
template <int S, int ALG>
__device__ __forceinline__ Color4 fusedCompositeInvocation(const FrameParams& p, const SrgbTables& tb, FusedArrays& A, int t, const AbufView& av,
                                                           size_t pix, int sampleID)
{
  const size_t P = av.viewSize;
  const int    L = p.L;
  const size_t ai = (size_t)sampleID * P + pix;
  switch(ALG)
  {
    case OIT_SIMPLE:
    case OIT_SPINLOCK:
    case OIT_INTERLOCK: {
      const size_t listPos = P * L * sampleID + pix;
      const int    n       = (int)min((uint32_t)L, av.aux[ai]);
      for(int i = 0; i < n; i++)
      {
        if(p.coverage)
        {
          const uint4 e = reinterpret_cast<const uint4*>(av.abuf)[listPos + (size_t)i * P];
          A.c[i][t] = e.x; A.d[i][t] = e.y; A.m[i][t] = e.z;
        }
        else
        {
          const uint2 e = reinterpret_cast<const uint2*>(av.abuf)[listPos + (size_t)i * P];
          A.c[i][t] = e.x; A.d[i][t] = e.y; A.m[i][t] = 0u;
        }
      }
      fusedBubbleSort(A, t, n);
      return fusedBlend<S>(tb, A, t, n, p.coverage != 0);
    }
    case OIT_LINKEDLIST: {
      const uint4* nodes  = reinterpret_cast<const uint4*>(av.abuf);
      uint32_t     offset = av.aux[ai];
      int          n      = 0;
      while(offset != 0u && n < L)
      {
        const uint4 e = nodes[offset];
        A.c[n][t] = e.x; A.d[n][t] = e.y; A.m[n][t] = e.z;
        n++;
        offset = e.w;
      }
      fusedBubbleSort(A, t, n);
      Color4 tailColor = zeroColor();
      while(offset != 0u)
      {
        const uint4 e = nodes[offset];
        if(p.tailBlend)
          doBlendPacked(tb, tailColor, fusedInsert<true>(A, t, L, e.x, e.y, e.z));
        else
          fusedInsert<false>(A, t, L, e.x, e.y, e.z);
        offset = e.w;
      }
      Color4 out = fusedBlend<S>(tb, A, t, n, p.coverage != 0);
      doBlend(out, tailColor);
      return out;
    }
    case OIT_LOOP: {
      const uint32_t* list = av.abuf + P * L * 2 * sampleID + pix;
      int             n    = 0;
      for(int i = 0; i < L; i++)
      {
        if(list[(size_t)i * P] == 0xFFFFFFFFu)
          break;
        n++;
      }
      list += P * L;
      Color4 out = zeroColor();
      for(int i = 0; i < n; i++)
        doBlendPacked(tb, out, list[(size_t)i * P]);
      return out;
    }
    case OIT_LOOP64: {
      const uint2* list = reinterpret_cast<const uint2*>(av.abuf) + P * L * sampleID + pix;
      Color4       out  = zeroColor();
      for(int i = 0; i < L; i++)
      {
        const uint2 e = list[(size_t)i * P];
        if(e.y == 0xFFFFFFFFu)
          break;
        doBlendPacked(tb, out, e.x);
      }
      return out;
    }
  }
  return zeroColor();
}

// composite (+ its ROP onto the shared-memory colour tile) of the tile pixel owned by thread t
// wAcc / wRev: the pixel's S WBOIT accumulator / revealage samples (shared-memory tile), only used for OIT_WEIGHTED
template <int S, int ALG>
__device__ __forceinline__ void fusedCompositePixel(const FrameParams& p, const SrgbTables& tb, FusedArrays& A, int t, const AbufView& av,
                                                    size_t pixA, size_t pix, uint32_t* px, const uint2* wAcc = nullptr,
                                                    const uint16_t* wRev = nullptr)
{
  if(ALG == OIT_WEIGHTED)
  {
    // K16 oitWeighted.frag.glsl:98-109 + BlendMode::WEIGHTED_COMPOSITE, per sample
#pragma unroll 1
    for(int s = 0; s < S; s++)
    {
      const size_t  idx = pix * S + s;
      const uint2   raw = wAcc ? wAcc[s] : reinterpret_cast<const uint2*>(p.wacc)[idx];
      const ushort4 acc = make_ushort4((unsigned short)(raw.x & 0xFFFFu), (unsigned short)(raw.x >> 16), (unsigned short)(raw.y & 0xFFFFu),
                                       (unsigned short)(raw.y >> 16));
      const float   a3  = h2f(acc.w);
      const float   den = a3 > 1e-5f ? a3 : 1e-5f;
      const Color4  src{__fdiv_rn(h2f(acc.x), den), __fdiv_rn(h2f(acc.y), den), __fdiv_rn(h2f(acc.z), den), h2f(wRev ? wRev[s] : p.wrev[idx])};
      px[s] = ropWeightedComposite(tb, px[s], src);
    }
    return;
  }
  if(p.sampleShading)
  {
#pragma unroll 1
    for(int s = 0; s < S; s++)
    {
      const Color4 out = fusedCompositeInvocation<S, ALG>(p, tb, A, t, av, pixA, s);
      if(!isZero(out))
        px[s] = ropPremult(tb, px[s], out);
    }
    return;
  }
  const Color4 out = fusedCompositeInvocation<S, ALG>(p, tb, A, t, av, pixA, 0);
  if(isZero(out))
    return;
  uint32_t prevDst = px[0], prevRes = ropPremult(tb, prevDst, out);
  px[0]            = prevRes;
#pragma unroll 1
  for(int s = 1; s < S; s++)
  {
    const uint32_t d = px[s];
    if(d != prevDst)
    {
      prevDst = d;
      prevRes = ropPremult(tb, d, out);
    }
    px[s] = prevRes;
  }
}

// one resolved pixel (gx, gy: output coordinates inside this band's rows) to m_viewportImage -- and, in split-frame mode
// over peer memory, straight into every band's whole-frame buffer (NVLink stores)
__device__ __forceinline__ void fusedStorePixel(const FrameParams& p, int gx, int gy, int outW, int ss, uint32_t result)
{
  p.fin[(size_t)gy * outW + gx] = result;
  if(p.peers)
  {
    const int    stripRows = p.stripTileRows * TILE_H / ss;
    const int    strip     = gy / stripRows;
    const size_t o         = (size_t)((strip * p.bandCount + p.bandIndex) * stripRows + (gy - strip * stripRows)) * outW + gx;
    for(int b = 0; b < p.bandCount; b++)
      p.peers->frame[b][o] = result;
  }
}

// copyOffscreenToBackBuffer for the tile: box resolve of the S samples / the ss x ss block, from the shared-memory tile
template <int S>
__device__ __forceinline__ void fusedResolveTile(const FrameParams& p, const SrgbTables& tb, const uint32_t* tileColor, int tileX0, int yLocal0,
                                                 int tid)
{
  const int ss   = p.supersample;
  const int outW = p.W / ss;
  const int side = TILE_W / ss;  // output pixels per tile side
  for(int o = tid; o < side * side; o += blockDim.x)
  {
    const int oy = o / side, ox = o - oy * side;
    const int gx = tileX0 / ss + ox, gy = yLocal0 / ss + oy;
    if(gx >= outW || gy >= p.localH / ss)
      continue;
    uint32_t result;
    // every sample of the block holds the same code (nothing but full-coverage blends touched the pixel: the common case):
    // the box filter of identical values is that value -- encode(decode(v) * (1 +- 1e-6)) == v for every 8-bit v, the
    // thresholds sit half a code away -- so the S decodes and the encode are skipped
    bool uniform = true;
    if(S == 1 && ss == 1)
      result = tileColor[oy * TILE_W + ox];
    else
    {
      result = tileColor[((oy * ss) * TILE_W + ox * ss) * S];
      for(int dy = 0; dy < ss; dy++)
        for(int dx = 0; dx < ss; dx++)
        {
          const uint32_t* px = tileColor + ((oy * ss + dy) * TILE_W + (ox * ss + dx)) * S;
          if(S % 4 == 0)
          {
#pragma unroll
            for(int s4 = 0; s4 < S / 4; s4++)
            {
              const uint4 v = reinterpret_cast<const uint4*>(px)[s4];
              uniform       = uniform && v.x == result && v.y == result && v.z == result && v.w == result;
            }
          }
          else
            uniform = uniform && px[0] == result;
        }
    }
    if(!uniform)
    {
      float sum[4] = {0.f, 0.f, 0.f, 0.f};
      for(int dy = 0; dy < ss; dy++)
        for(int dx = 0; dx < ss; dx++)
        {
          const uint32_t* px = tileColor + ((oy * ss + dy) * TILE_W + (ox * ss + dx)) * S;
#pragma unroll
          for(int s = 0; s < S; s++)
          {
            const Color4 d = decodeDst(tb, px[s]);
            sum[0]         = __fadd_rn(sum[0], d.r);
            sum[1]         = __fadd_rn(sum[1], d.g);
            sum[2]         = __fadd_rn(sum[2], d.b);
            sum[3]         = __fadd_rn(sum[3], d.a);
          }
        }
      const float inv = 1.0f / (float)(ss * ss * S);
      result = encodeDst(tb, Color4{__fmul_rn(sum[0], inv), __fmul_rn(sum[1], inv), __fmul_rn(sum[2], inv), __fmul_rn(sum[3], inv)});
    }
    fusedStorePixel(p, gx, gy, outW, ss, result);
  }
}

// a tile nothing was drawn into (and no opaque pass ran): every sample holds the clear colour, which resolves to itself
__device__ __forceinline__ void fusedClearTile(const FrameParams& p, int tileX0, int yLocal0, int tid)
{
  const int ss   = p.supersample;
  const int outW = p.W / ss;
  const int side = TILE_W / ss;
  for(int o = tid; o < side * side; o += blockDim.x)
  {
    const int oy = o / side, ox = o - oy * side;
    const int gx = tileX0 / ss + ox, gy = yLocal0 / ss + oy;
    if(gx < outW && gy < p.localH / ss)
      fusedStorePixel(p, gx, gy, outW, ss, p.clearColor);
  }
}

}  // namespace oit
