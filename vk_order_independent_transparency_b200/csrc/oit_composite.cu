// oit_composite.cu -- clears, the per-pixel sort + composite programs, and the resolve.
//
// Replaces:
//   F1  clearTransparent{Simple,LinkedList,Loop,Loop64,Lock} + render-pass clears   oitRender.cpp:156-356, :89-91, :394-397
//   K3/K12/K14 oitSimple.frag.glsl:115-169 (== oitSpinlock:153-207 == oitInterlock:176-230)
//   K5  oitLinkedList.frag.glsl:106-172      K8  oitLoop.frag.glsl:191-222      K10 oitLoop64.frag.glsl:162-183
//   K16 oitWeighted.frag.glsl:98-109         sorts oitCompositeDefines.glsl:51-139
//   F3  ROP blend of the composite output (main.cpp:548-558, :576-588)              F4 copyOffscreenToBackBuffer main.cpp:645-774
//
// The composite is one thread per pixel (per sample when sample shading): the <= OIT_LAYERS fragments live in
// registers (fully unrolled, statically indexed arrays) and the reference's bubble sort is executed as the fixed
// compare-exchange network it is, so ties resolve exactly as in the GLSL.
#include "oit_device.cuh"

namespace oit {

// ---- fills -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_fill32(uint32_t* __restrict__ dst, size_t nWords, uint32_t value)
{
  const size_t n4     = nWords / 4;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  uint4*       d4     = reinterpret_cast<uint4*>(dst);
  const uint4  v4     = make_uint4(value, value, value, value);
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride)
    d4[i] = v4;
  for(size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nWords; i += stride)
    dst[i] = value;
}
static int fill32(void* dst, size_t nWords, uint32_t value, cudaStream_t s)
{
  if(dst == nullptr || nWords == 0)
    return 0;
  const size_t want = (nWords / 4 + 255) / 256;
  const int    grid = (int)(want < 1 ? 1 : (want > 148 * 8 ? 148 * 8 : want));
  k_fill32<<<grid, 256, 0, s>>>(reinterpret_cast<uint32_t*>(dst), nWords, value);
  return 1;
}

int launchClears(const FrameParams& p, int algorithm, cudaStream_t s, bool skipListHeads, bool skipCounter)
{
  const size_t P        = (size_t)p.W * p.localH;
  const size_t auxWords = P * p.layers;
  int          n        = 0;
  switch(p.onChip ? -1 : algorithm)  // on chip: the fused frame kernel clears its shared-memory slice itself
  {
    case OIT_SIMPLE: n += fill32(p.aux, auxWords, 0u, s); break;
    case OIT_LINKEDLIST:
      if(!skipListHeads)
        n += fill32(p.aux, auxWords, 0u, s);
      if(!skipCounter)
        n += fill32(p.counter, 1, 0u, s);
      break;
    case OIT_LOOP:
      // only the depth half of every sample block (oitRender.cpp:252-261)
      for(int i = 0; i < p.layers; i++)
        n += fill32(p.abuf + (size_t)i * P * p.L * 2, P * p.L, 0xFFFFFFFFu, s);
      break;
    case OIT_LOOP64: n += fill32(p.abuf, P * p.L * 2 * p.layers, 0xFFFFFFFFu, s); break;
    case OIT_SPINLOCK:
    case OIT_INTERLOCK:
      n += fill32(p.adepth, auxWords, 0xFFFFFFFFu, s);
      n += fill32(p.aux, auxWords, 0u, s);
      if(algorithm == OIT_SPINLOCK)
        n += fill32(p.spin, auxWords, 0u, s);
      break;
    case OIT_WEIGHTED:
      if(!p.fused)  // the fused frame kernel keeps the WBOIT targets of a tile in shared memory
      {
        n += fill32(p.wacc, P * p.msaa * 2, 0u, s);               // RGBA16F (0,0,0,0)
        n += fill32(p.wrev, (P * p.msaa + 1) / 2, 0x3C003C00u, s);  // R16F 1.0
      }
      break;
  }
  if(!(p.fused && p.depth == nullptr))  // the fused frame kernel keeps the colour tile in shared memory
    n += fill32(p.color, P * p.msaa, p.clearColor, s);
  if(p.depth)
    n += fill32(p.depth, P * p.msaa, 0x3F800000u, s);
  return n;
}

// ---- sorted composites (Simple / Spinlock / Interlock / Linked List) --------------------------------------------------
struct Elem
{
  uint32_t c, d, m;
};
__device__ __forceinline__ bool depthGE(uint32_t a, uint32_t b) { return __uint_as_float(a) >= __uint_as_float(b); }
__device__ __forceinline__ bool depthLT(uint32_t a, uint32_t b) { return __uint_as_float(a) < __uint_as_float(b); }

// bubbleSort(array, n) of oitCompositeDefines.glsl:51-89 as its compare-exchange network (swap on >=)
template <int LMAX>
__device__ __forceinline__ void bubbleSort(Elem (&a)[LMAX], int n)
{
#pragma unroll
  for(int i = LMAX - 2; i >= 0; --i)
  {
    if(i <= n - 2)
    {
#pragma unroll
      for(int j = 0; j <= i; ++j)
      {
        if(depthGE(a[j].d, a[j + 1].d))
        {
          const Elem t = a[j + 1];
          a[j + 1]     = a[j];
          a[j]         = t;
        }
      }
    }
  }
}

// insertionSortTail / insertionSort (oitCompositeDefines.glsl:94-139) on the first L entries; returns what falls out
template <int LMAX, bool TAIL>
__device__ __forceinline__ Elem insertSorted(Elem (&a)[LMAX], int L, const Elem item)
{
  Elem last = a[0];
#pragma unroll
  for(int i = 0; i < LMAX; i++)
    if(i == L - 1)
      last = a[i];
  Elem newlast = item;
  if(!TAIL || depthLT(item.d, last.d))
  {
    int pos = L;
#pragma unroll
    for(int i = LMAX - 1; i >= 0; i--)
      if(i < L && depthLT(item.d, a[i].d))
        pos = i;
    if(pos < L)
    {
      newlast = last;
#pragma unroll
      for(int j = LMAX - 1; j >= 1; j--)
        if(j < L && j > pos)
          a[j] = a[j - 1];
#pragma unroll
      for(int j = 0; j < LMAX; j++)
        if(j == pos)
          a[j] = item;
    }
  }
  return newlast;
}

// blend of the sorted fragments (oitSimple.frag.glsl:138-167): per-sample coverage loop, or plain front-to-back
template <int LMAX, int COV>
__device__ __forceinline__ Color4 blendSorted(const SrgbTables& t, const Elem (&a)[LMAX], int n)
{
  Color4 sum = zeroColor();
  if(COV > 1 && LMAX <= 16)
  {
    // unpack + premultiply every fragment once (doBlendPacked is a pure function of the packed word), then run the
    // per-sample coverage loops on registers
    Color4 pm[LMAX];
#pragma unroll
    for(int i = 0; i < LMAX; i++)
      if(i < n)
        pm[i] = premultiply(unpackColor(t, a[i].c));
#pragma unroll 1
    for(int s = 0; s < COV; s++)
    {
      Color4 sc = zeroColor();
#pragma unroll
      for(int i = 0; i < LMAX; i++)
        if(i < n && (a[i].m & (1u << s)))
          doBlend(sc, pm[i]);
      sum.r = __fadd_rn(sum.r, sc.r);
      sum.g = __fadd_rn(sum.g, sc.g);
      sum.b = __fadd_rn(sum.b, sc.b);
      sum.a = __fadd_rn(sum.a, sc.a);
    }
    const float inv = 1.0f / (float)COV;
    sum.r           = __fmul_rn(sum.r, inv);
    sum.g           = __fmul_rn(sum.g, inv);
    sum.b           = __fmul_rn(sum.b, inv);
    sum.a           = __fmul_rn(sum.a, inv);
  }
  else if(COV > 1)
  {
#pragma unroll 1
    for(int s = 0; s < COV; s++)
    {
      Color4 sc = zeroColor();
#pragma unroll
      for(int i = 0; i < LMAX; i++)
        if(i < n && (a[i].m & (1u << s)))
          doBlendPacked(t, sc, a[i].c);
      sum.r = __fadd_rn(sum.r, sc.r);
      sum.g = __fadd_rn(sum.g, sc.g);
      sum.b = __fadd_rn(sum.b, sc.b);
      sum.a = __fadd_rn(sum.a, sc.a);
    }
    const float inv = 1.0f / (float)COV;
    sum.r           = __fmul_rn(sum.r, inv);
    sum.g           = __fmul_rn(sum.g, inv);
    sum.b           = __fmul_rn(sum.b, inv);
    sum.a           = __fmul_rn(sum.a, inv);
  }
  else
  {
#pragma unroll
    for(int i = 0; i < LMAX; i++)
      if(i < n)
        doBlendPacked(t, sum, a[i].c);
  }
  return sum;
}

// ROP write of a composite output: onto every sample of the pixel, or onto one sample when sample shading
__device__ __forceinline__ void ropComposite(const FrameParams& p, const SrgbTables& t, size_t pix, int sampleID, bool perSample,
                                             const Color4& out)
{
  if(isZero(out))
    return;
  uint32_t* px = p.color + pix * p.msaa;
  if(perSample)
    px[sampleID] = ropPremult(t, px[sampleID], out);
  else
  {
    // the blend is a pure function of (dst, src): samples holding the same destination word share one evaluation
    uint32_t prevDst = px[0], prevRes = ropPremult(t, prevDst, out);
    px[0]            = prevRes;
    for(int s = 1; s < p.msaa; s++)
    {
      const uint32_t d = px[s];
      if(d != prevDst)
      {
        prevDst = d;
        prevRes = ropPremult(t, d, out);
      }
      px[s] = prevRes;
    }
  }
}

// KIND 0: fixed-slot k-buffer (Simple / Spinlock / Interlock); KIND 1: linked list
template <int LMAX, int COV, int KIND>
__global__ void __launch_bounds__(128) k_composite_sorted(const FrameParams p)
{
  __shared__ SrgbTables tabs;
  loadTables(tabs, p.tables);
  __syncthreads();
  const size_t P     = (size_t)p.W * p.localH;
  const size_t total = P * p.layers;
  const size_t idx   = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= total)
    return;
  const int    sampleID = (int)(idx / P);
  const size_t pix      = idx - (size_t)sampleID * P;
  const int    L        = p.L;
  Elem         arr[LMAX];
  int          n = 0;
  Color4       out;
  if(KIND == 0)
  {
    const size_t listPos = P * L * sampleID + pix;
    n                    = (int)min((uint32_t)L, p.aux[idx]);
#pragma unroll
    for(int i = 0; i < LMAX; i++)
      if(i < n)
      {
        if(COV > 1)
        {
          const uint4 e = reinterpret_cast<const uint4*>(p.abuf)[listPos + (size_t)i * P];
          arr[i]        = Elem{e.x, e.y, e.z};
        }
        else
        {
          const uint2 e = reinterpret_cast<const uint2*>(p.abuf)[listPos + (size_t)i * P];
          arr[i]        = Elem{e.x, e.y, 0u};
        }
      }
    bubbleSort<LMAX>(arr, n);
    out = blendSorted<LMAX, COV>(tabs, arr, n);
  }
  else
  {
    uint32_t     offset = p.aux[idx];
    const uint4* nodes  = reinterpret_cast<const uint4*>(p.abuf);
#pragma unroll
    for(int i = 0; i < LMAX; i++)
      if(offset != 0u && offset < p.capacity && i < L)
      {
        const uint4 e = nodes[offset];
        arr[i]        = Elem{e.x, e.y, e.z};
        n             = i + 1;
        offset        = e.w;
      }
    bubbleSort<LMAX>(arr, n);
    Color4 tailColor = zeroColor();
    // (a list cannot be longer than the pool: the bound keeps a corrupt A-buffer -- oit_upload of a foreign dump -- from
    // spinning forever in a cycle; an index outside the pool ends the list)
    for(uint32_t guard = p.capacity; offset != 0u && offset < p.capacity && guard != 0u; guard--)
    {
      const uint4 e = nodes[offset];
      const Elem  it{e.x, e.y, e.z};
      if(p.tailBlend)
      {
        const Elem tail = insertSorted<LMAX, true>(arr, L, it);
        doBlendPacked(tabs, tailColor, tail.c);
      }
      else
        insertSorted<LMAX, false>(arr, L, it);
      offset = e.w;
    }
    out = blendSorted<LMAX, COV>(tabs, arr, n);
    doBlend(out, tailColor);
  }
  ropComposite(p, tabs, pix, sampleID, p.sampleShading != 0, out);
}

// K8 oitLoop.frag.glsl:191-222 and K10 oitLoop64.frag.glsl:162-183: the slots are already sorted
template <bool LOOP64>
__global__ void __launch_bounds__(256) k_composite_loop(const FrameParams p)
{
  __shared__ SrgbTables tabs;
  loadTables(tabs, p.tables);
  __syncthreads();
  const size_t P     = (size_t)p.W * p.localH;
  const size_t total = P * p.layers;
  const size_t idx   = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= total)
    return;
  const int    sampleID = (int)(idx / P);
  const size_t pix      = idx - (size_t)sampleID * P;
  const int    L        = p.L;
  Color4       out      = zeroColor();
  if(LOOP64)
  {
    const uint2* list = reinterpret_cast<const uint2*>(p.abuf) + P * L * sampleID + pix;
    for(int i = 0; i < L; i++)
    {
      const uint2 e = list[(size_t)i * P];
      if(e.y == 0xFFFFFFFFu)
        break;
      doBlendPacked(tabs, out, e.x);
    }
  }
  else
  {
    const uint32_t* list = p.abuf + P * L * 2 * sampleID + pix;
    int             n    = 0;
    for(int i = 0; i < L; i++)
    {
      if(list[(size_t)i * P] == 0xFFFFFFFFu)
        break;
      n++;
    }
    list += P * L;
    for(int i = 0; i < n; i++)
      doBlendPacked(tabs, out, list[(size_t)i * P]);
  }
  ropComposite(p, tabs, pix, sampleID, p.sampleShading != 0, out);
}

// K16 oitWeighted.frag.glsl:98-109 + BlendMode::WEIGHTED_COMPOSITE; one thread per colour sample
__global__ void __launch_bounds__(256) k_composite_weighted(const FrameParams p)
{
  __shared__ SrgbTables tabs;
  loadTables(tabs, p.tables);
  __syncthreads();
  const size_t total = (size_t)p.W * p.localH * p.msaa;
  const size_t idx   = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= total)
    return;
  const ushort4 acc = reinterpret_cast<const ushort4*>(p.wacc)[idx];
  const float   a3  = h2f(acc.w);
  const float   den = a3 > 1e-5f ? a3 : 1e-5f;
  const Color4  src{__fdiv_rn(h2f(acc.x), den), __fdiv_rn(h2f(acc.y), den), __fdiv_rn(h2f(acc.z), den), h2f(p.wrev[idx])};
  p.color[idx] = ropWeightedComposite(tabs, p.color[idx], src);
}

template <int KIND>
static void launchSorted(const FrameParams& p, unsigned grid, cudaStream_t s)
{
  const int cov = p.coverage ? p.msaa : 1;
#define OIT_SORTED(LM)                                                                                                           \
  do                                                                                                                             \
  {                                                                                                                              \
    if(cov == 1)                                                                                                                 \
      k_composite_sorted<LM, 1, KIND><<<grid, 128, 0, s>>>(p);                                                                   \
    else if(cov == 4)                                                                                                            \
      k_composite_sorted<LM, 4, KIND><<<grid, 128, 0, s>>>(p);                                                                   \
    else                                                                                                                         \
      k_composite_sorted<LM, 8, KIND><<<grid, 128, 0, s>>>(p);                                                                   \
  } while(0)
  if(p.L <= 1)
    OIT_SORTED(1);
  else if(p.L <= 2)
    OIT_SORTED(2);
  else if(p.L <= 4)
    OIT_SORTED(4);
  else if(p.L <= 8)
    OIT_SORTED(8);
  else if(p.L <= 16)
    OIT_SORTED(16);
  else
    OIT_SORTED(32);
#undef OIT_SORTED
}

int launchComposite(const FrameParams& p, int algorithm, cudaStream_t s)
{
  const size_t P     = (size_t)p.W * p.localH;
  const size_t total = P * p.layers;
  if(P == 0)
    return 0;
  switch(algorithm)
  {
    case OIT_SIMPLE:
    case OIT_SPINLOCK:
    case OIT_INTERLOCK: launchSorted<0>(p, (unsigned)((total + 127) / 128), s); break;
    case OIT_LINKEDLIST: launchSorted<1>(p, (unsigned)((total + 127) / 128), s); break;
    case OIT_LOOP: k_composite_loop<false><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p); break;
    case OIT_LOOP64: k_composite_loop<true><<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p); break;
    case OIT_WEIGHTED: k_composite_weighted<<<(unsigned)((P * p.msaa + 255) / 256), 256, 0, s>>>(p); break;
    default: return 0;
  }
  return 1;
}

// ---- resolve: copyOffscreenToBackBuffer (main.cpp:645-774) --------------------------------------------------------------
// box average of the msaa samples (vkCmdResolveImage) or of the ss x ss block (vkCmdBlitImage, LINEAR, exact 2x), done in
// linear space on the sRGB target, then the raw copy that keeps the sRGB-encoded bytes.
__global__ void __launch_bounds__(256) k_resolve(const FrameParams p, int ss, int outW, int outLocalH)
{
  __shared__ SrgbTables tabs;
  loadTables(tabs, p.tables);
  __syncthreads();
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= (size_t)outW * outLocalH)
    return;
  const int y = (int)(idx / outW), x = (int)(idx - (size_t)y * outW);
  const int S = p.msaa;
  if(S == 1 && ss == 1)
  {
    p.fin[idx] = p.color[idx];
    return;
  }
  float sum[4] = {0.f, 0.f, 0.f, 0.f};
  for(int dy = 0; dy < ss; dy++)
    for(int dx = 0; dx < ss; dx++)
    {
      const uint32_t* px = p.color + ((size_t)(y * ss + dy) * p.W + (x * ss + dx)) * S;
      for(int s = 0; s < S; s++)
      {
        const Color4 d = decodeDst(tabs, px[s]);
        sum[0]         = __fadd_rn(sum[0], d.r);
        sum[1]         = __fadd_rn(sum[1], d.g);
        sum[2]         = __fadd_rn(sum[2], d.b);
        sum[3]         = __fadd_rn(sum[3], d.a);
      }
    }
  const float inv = 1.0f / (float)(ss * ss * S);
  p.fin[idx]      = encodeDst(tabs, Color4{__fmul_rn(sum[0], inv), __fmul_rn(sum[1], inv), __fmul_rn(sum[2], inv), __fmul_rn(sum[3], inv)});
}

int launchResolve(const FrameParams& p, int supersample, int outW, int outLocalH, cudaStream_t s)
{
  const size_t n = (size_t)outW * outLocalH;
  if(n == 0)
    return 0;
  k_resolve<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p, supersample, outW, outLocalH);
  return 1;
}

}  // namespace oit
