// oit_fused.cuh -- composite + resolve of ONE tile by the CTA that has just finished the tile's colour pass.
//
// Same programs as oit_composite.cu (K3/K5/K8/K10/K12/K14/K16 + copyOffscreenToBackBuffer, cited there), written for
// OIT_LAYERS <= 8: the depths of a pixel's fragments are sorted in REGISTERS (a compare-exchange network that replays the
// reference's bubble sort, and a select-chain insertion for the tail), together with 4-bit slot numbers that say where the
// fragment's colour and coverage mask sit in shared memory (row `slot` of thread t's column: conflict free).
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

// What stays in shared memory per pixel: colour and coverage mask of its <= 8 sorted fragments, addressed through a slot
// number.  The DEPTHS and the slot numbers live in registers, where the sorting happens.
struct FusedArrays
{
  uint32_t c[FUSED_LCAP][TILE_PIX];
  uint32_t m[FUSED_LCAP][TILE_PIX];
};

// The sorted fragments of one composite invocation: depth d[k] and storage slot ix[k] (FusedArrays row) of the k-th entry.
struct SortRegs
{
  float    d[FUSED_LCAP];
  uint32_t ix[FUSED_LCAP];
};

// bubbleSort of oitCompositeDefines.glsl:51-89 on the first n entries, as the fixed compare-exchange network it is (swap on
// >=, so equal depths end up exactly where the reference's loop leaves them): pass i = n-2 .. 0 compares (j, j+1), j <= i.
// Entries are left-aligned; a pass the reference does not run (i > n-2) is predicated off, and passes that no lane of the
// warp needs are skipped (the pixels of a warp have lists of similar length, see the callers).
__device__ __forceinline__ void fusedSortNetwork(SortRegs& r, int n)
{
  const int nWarp = __reduce_max_sync(__activemask(), n);
#pragma unroll
  for(int i = FUSED_LCAP - 2; i >= 0; --i)
  {
    if(i > nWarp - 2)
      continue;
    const bool pass = i <= n - 2;
#pragma unroll
    for(int j = 0; j <= i; ++j)
    {
      const bool     sw = pass && r.d[j] >= r.d[j + 1];
      const float    dl = sw ? r.d[j + 1] : r.d[j], dh = sw ? r.d[j] : r.d[j + 1];
      const uint32_t il = sw ? r.ix[j + 1] : r.ix[j], ih = sw ? r.ix[j] : r.ix[j + 1];
      r.d[j] = dl; r.d[j + 1] = dh; r.ix[j] = il; r.ix[j + 1] = ih;
    }
  }
}

// insertionSortTail / insertionSort (oitCompositeDefines.glsl:94-139) of one more fragment into the L sorted entries: the
// new one goes in front of the first entry it is nearer than, the last entry falls out (its slot takes the new colour and
// mask).  Returns the colour that falls out -- the new fragment's own if it is not nearer than the last entry.
__device__ __forceinline__ uint32_t fusedInsertRegs(FusedArrays& A, int t, SortRegs& r, int L, uint32_t c, float d, uint32_t m)
{
  // entry L-1 (a register picked by a run-time index: a select chain; L == 8 in the default configuration)
  float    dLast = r.d[FUSED_LCAP - 1];
  uint32_t iLast = r.ix[FUSED_LCAP - 1];
#pragma unroll
  for(int j = 0; j < FUSED_LCAP - 1; j++)
    if(j == L - 1)
    {
      dLast = r.d[j];
      iLast = r.ix[j];
    }
  if(!(d < dLast))
    return c;
  const uint32_t outColor = A.c[iLast][t];
  A.c[iLast][t]           = c;
  A.m[iLast][t]           = m;
  bool lt[FUSED_LCAP];
#pragma unroll
  for(int j = 0; j < FUSED_LCAP; j++)
    lt[j] = d < r.d[j];
#pragma unroll
  for(int j = FUSED_LCAP - 1; j >= 1; j--)
    if(j < L)
    {
      r.d[j]  = lt[j - 1] ? r.d[j - 1] : (lt[j] ? d : r.d[j]);
      r.ix[j] = lt[j - 1] ? r.ix[j - 1] : (lt[j] ? iLast : r.ix[j]);
    }
  r.d[0]  = lt[0] ? d : r.d[0];
  r.ix[0] = lt[0] ? iLast : r.ix[0];
  return outColor;
}

// slot numbers of the sorted entries, 4 bits each (the blend loop is not unrolled over the fragments)
__device__ __forceinline__ uint32_t fusedPackSlots(const SortRegs& r)
{
  uint32_t packed = 0u;
#pragma unroll
  for(int k = 0; k < FUSED_LCAP; k++)
    packed |= r.ix[k] << (4 * k);
  return packed;
}

// blend of the sorted fragments (oitSimple.frag.glsl:138-167).  Coverage mode keeps one accumulator per sample and walks
// the fragments once: every sample still sees its fragments front to back, so the arithmetic is the reference's.
template <int S>
__device__ __forceinline__ Color4 fusedBlend(const SrgbTables& tb, const FusedArrays& A, int t, int n, uint32_t slots, bool coverage)
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
      const uint32_t slot = (slots >> (4 * i)) & 15u;
      const Color4   pm   = premultiply(unpackColor(tb, A.c[slot][t]));
      const uint32_t m    = A.m[slot][t];
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
    doBlendPacked(tb, sum, A.c[(slots >> (4 * i)) & 15u][t]);
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
      SortRegs     r;
#pragma unroll
      for(int i = 0; i < FUSED_LCAP; i++)
      {
        r.d[i]  = 0.f;
        r.ix[i] = (uint32_t)i;
        if(i < n)
        {
          if(p.coverage)
          {
            const uint4 e = reinterpret_cast<const uint4*>(av.abuf)[listPos + (size_t)i * P];
            A.c[i][t] = e.x; r.d[i] = __uint_as_float(e.y); A.m[i][t] = e.z;
          }
          else
          {
            const uint2 e = reinterpret_cast<const uint2*>(av.abuf)[listPos + (size_t)i * P];
            A.c[i][t] = e.x; r.d[i] = __uint_as_float(e.y); A.m[i][t] = 0u;
          }
        }
      }
      fusedSortNetwork(r, n);
      return fusedBlend<S>(tb, A, t, n, fusedPackSlots(r), p.coverage != 0);
    }
    case OIT_LINKEDLIST: {
      const uint4* nodes  = reinterpret_cast<const uint4*>(av.abuf);
      uint32_t     offset = av.aux[ai];
      int          n      = 0;
      SortRegs     r;
#ifdef OIT_EXPERIMENT_NO_CHASE
      uint32_t remaining = (uint32_t)av.viewSize;  // timing experiment: the list is walked as if its nodes were contiguous
      if(offset < remaining)
        remaining = offset;
#endif
#pragma unroll
      for(int i = 0; i < FUSED_LCAP; i++)
      {
        r.d[i]  = 0.f;
        r.ix[i] = (uint32_t)i;
        if(offset != 0u && offset < p.capacity && i < L)
        {
          const uint4 e = nodes[offset];
          A.c[i][t] = e.x; r.d[i] = __uint_as_float(e.y); A.m[i][t] = e.z;
          n      = i + 1;
#ifdef OIT_EXPERIMENT_NO_CHASE
          offset = --remaining ? offset - 1u : 0u;
#else
          offset = e.w;
#endif
        }
      }
      fusedSortNetwork(r, n);
      Color4 tailColor = zeroColor();
      // (a list cannot be longer than the pool: the bound keeps a corrupt A-buffer from spinning forever in a cycle)
      for(uint32_t guard = p.capacity; offset != 0u && offset < p.capacity && guard != 0u; guard--)
      {
        const uint4    e   = nodes[offset];
        const uint32_t out = fusedInsertRegs(A, t, r, L, e.x, __uint_as_float(e.y), e.z);
        if(p.tailBlend)
          doBlendPacked(tb, tailColor, out);
#ifdef OIT_EXPERIMENT_NO_CHASE
        offset = --remaining ? offset - 1u : 0u;
#else
        offset = e.w;
#endif
      }
      Color4 out = fusedBlend<S>(tb, A, t, n, fusedPackSlots(r), p.coverage != 0);
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
  if(p.peers && p.pushers == 0)
  {
    const int    stripRows = p.stripTileRows * TILE_H / ss;
    const int    strip     = gy / stripRows;
    const size_t o         = (size_t)((strip * p.bandCount + p.bandIndex) * stripRows + (gy - strip * stripRows)) * outW + gx;
    uint32_t* const* frames = peerFramesOfRunningFrame(p.peers, p.bandIndex);
    for(int b = 0; b < p.bandCount; b++)
      frames[b][o] = result;
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

// Composite + ROP + resolve of ONE pixel whose samples nobody else touches any more (coverage shading or no AA, no
// super-sampling: the output pixel is the pixel): the S destination samples are read once, blended and box-filtered in
// registers, and only the resolved pixel is stored -- no barrier between the composite and the resolve of a tile.
// `covered`: the pixel has fragments to composite (otherwise only the resolve of what the opaque pass / clear left).
template <int S, int ALG>
__device__ __forceinline__ void fusedFinishPixel(const FrameParams& p, const SrgbTables& tb, FusedArrays& A, int t, const AbufView& av, size_t pixA,
                                                 const uint32_t* px, bool covered, int gx, int gyLocal)
{
  uint32_t v[S];
  if(S % 4 == 0)
  {
#pragma unroll
    for(int q = 0; q < S / 4; q++)
    {
      const uint4 w = reinterpret_cast<const uint4*>(px)[q];
      v[4 * q] = w.x; v[4 * q + 1] = w.y; v[4 * q + 2] = w.z; v[4 * q + 3] = w.w;
    }
  }
  else
  {
#pragma unroll
    for(int s = 0; s < S; s++)
      v[s] = px[s];
  }
  if(covered)
  {
    const Color4 out = fusedCompositeInvocation<S, ALG>(p, tb, A, t, av, pixA, 0);
    if(!isZero(out))
    {
      // the same source colour goes to every sample: samples that hold the same destination word share the blend
      uint32_t prevDst = v[0], prevRes = ropPremult(tb, prevDst, out);
      v[0]             = prevRes;
#pragma unroll
      for(int s = 1; s < S; s++)
      {
        if(v[s] != prevDst)
        {
          prevDst = v[s];
          prevRes = ropPremult(tb, prevDst, out);
        }
        v[s] = prevRes;
      }
    }
  }
  uint32_t result  = v[0];
  bool     uniform = true;
#pragma unroll
  for(int s = 1; s < S; s++)
    uniform = uniform && v[s] == result;
  if(!uniform)
  {
    // (the box filter of identical codes is that code, see fusedResolveTile)
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for(int s = 0; s < S; s++)
    {
      const Color4 d = decodeDst(tb, v[s]);
      sum[0]         = __fadd_rn(sum[0], d.r);
      sum[1]         = __fadd_rn(sum[1], d.g);
      sum[2]         = __fadd_rn(sum[2], d.b);
      sum[3]         = __fadd_rn(sum[3], d.a);
    }
    const float inv = 1.0f / (float)S;
    result = encodeDst(tb, Color4{__fmul_rn(sum[0], inv), __fmul_rn(sum[1], inv), __fmul_rn(sum[2], inv), __fmul_rn(sum[3], inv)});
  }
  fusedStorePixel(p, gx, gyLocal, p.W, 1, result);
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
