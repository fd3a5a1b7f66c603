// oit_raster_common.cuh -- pieces shared by the tile raster kernels (oit_raster.cu, oit_raster_ll.cu): sample patterns,
// fixed-point coverage with the top-left rule, the early per-sample depth test, and the per-triangle set-up.
// Reference: fixed-function raster / early depth / post-depth coverage, main.cpp:504-532, oitColorDepthDefines.glsl:36-37.
#pragma once
#include "oit_fragment.cuh"
#include "oit_fused.cuh"

namespace oit {

#ifndef OIT_ITEMS_PER_THREAD
#define OIT_ITEMS_PER_THREAD 4
#endif
#ifndef OIT_SMEM_CARVEOUT
#define OIT_SMEM_CARVEOUT -1
#endif
#ifndef OIT_USE_DP2A
#define OIT_USE_DP2A 1
#endif
#ifndef OIT_TICKET_MATCH
#define OIT_TICKET_MATCH 1
#endif
constexpr int ITEMS_PER_THREAD = OIT_ITEMS_PER_THREAD;
constexpr int BATCH_ITEMS      = RASTER_THREADS * ITEMS_PER_THREAD;
constexpr int MASK_WORDS       = RASTER_THREADS / 32;

template <int S>
struct SamplePattern;
template <>
struct SamplePattern<1>
{
  static __device__ __forceinline__ int x(int) { return 128; }
  static __device__ __forceinline__ int y(int) { return 128; }
  static __device__ __forceinline__ int xr(int) { return 128; }  // xr / yr: the same for a run-time sample index
  static __device__ __forceinline__ int yr(int) { return 128; }
};
template <>
struct SamplePattern<4>
{
  static __device__ __forceinline__ int x(int s) { return s == 0 ? 96 : s == 1 ? 224 : s == 2 ? 32 : 160; }
  static __device__ __forceinline__ int y(int s) { return s == 0 ? 32 : s == 1 ? 96 : s == 2 ? 160 : 224; }
  static __device__ __forceinline__ int xr(int s) { return (int)((0xA020E060u >> (8 * s)) & 255u); }
  static __device__ __forceinline__ int yr(int s) { return (int)((0xE0A06020u >> (8 * s)) & 255u); }
};
template <>
struct SamplePattern<8>
{
  static __device__ __forceinline__ int x(int s)
  {
    return s == 0 ? 144 : s == 1 ? 112 : s == 2 ? 208 : s == 3 ? 80 : s == 4 ? 48 : s == 5 ? 16 : s == 6 ? 176 : 240;
  }
  static __device__ __forceinline__ int y(int s)
  {
    return s == 0 ? 80 : s == 1 ? 176 : s == 2 ? 144 : s == 3 ? 48 : s == 4 ? 208 : s == 5 ? 112 : s == 6 ? 240 : 16;
  }
  static __device__ __forceinline__ int xr(int s) { return (int)((0xF0B0103050D07090ull >> (8 * s)) & 255ull); }
  static __device__ __forceinline__ int yr(int s) { return (int)((0x10F070D03090B050ull >> (8 * s)) & 255ull); }
};

// c + lo16(a) * byte0(b) + hi16(a) * byte1(b), a signed halves, b unsigned bytes (SASS IDP.2A.LO.S16.U8)
__device__ __forceinline__ int dp2aS16U8(int a, uint32_t b, int c)
{
  int d;
  asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// unbiased edge function q (from vertex q+1 to vertex q+2) at a point given in 1/256 px, as a float.
// SMALL triangles (extent <= 64 px) fit int32; the conversion to float rounds identically either way.
__device__ __forceinline__ float edgeFloat(const TriSlot& s, int q, int px, int py, bool small)
{
  const int a = (q + 1) % 3, b = (q + 2) % 3;
  const int dx = s.x[b] - s.x[a], dy = s.y[b] - s.y[a];
  if(small)
    return __int2float_rn(dx * (py - s.y[a]) - dy * (px - s.x[a]));
  return __ll2float_rn((long long)dx * (py - s.y[a]) - (long long)dy * (px - s.x[a]));
}

// post-depth coverage mask of pixel (gx, gy) [global], depth samples at dpx (or nullptr = cleared to 1.0)
// coverage of a triangle too large for the int32 / IDP.2A path (extent > 64 px): rare, kept out of line
template <int S>
__device__ OIT_COLD uint32_t coverageMaskLarge(const TriSlot& s, int ox, int oy)
{
  uint32_t  mask = 0;
  long long e[3];
  int       dxs[3], dys[3];
#pragma unroll
  for(int q = 0; q < 3; q++)
  {
    const int a = (q + 1) % 3, b = (q + 2) % 3;
    dxs[q]      = s.x[b] - s.x[a];
    dys[q]      = s.y[b] - s.y[a];
    e[q]        = (long long)dxs[q] * (oy - s.y[a]) - (long long)dys[q] * (ox - s.x[a]) - (long long)((s.box >> (16 + q)) & 1u);
  }
#pragma unroll 1
  for(int sI = 0; sI < S; sI++)
  {
    const int       sx = SamplePattern<S>::xr(sI), sy = SamplePattern<S>::yr(sI);
    const long long e0 = e[0] + (long long)dxs[0] * sy - (long long)dys[0] * sx;
    const long long e1 = e[1] + (long long)dxs[1] * sy - (long long)dys[1] * sx;
    const long long e2 = e[2] + (long long)dxs[2] * sy - (long long)dys[2] * sx;
    if((e0 | e1 | e2) >= 0)
      mask |= 1u << sI;
  }
  return mask;
}

// early per-sample depth test, VK_COMPARE_OP_LESS (main.cpp:530-532): only runs when opaque geometry was drawn or a vertex
// depth is not safely below the clear value 1.0
template <int S>
__device__ OIT_COLD uint32_t depthTestMask(const TriSlot& s, int ox, int oy, const float* dpx, uint32_t mask)
{
  const bool small = (s.box >> 20) & 1u;
#pragma unroll 1
  for(int sI = 0; sI < S; sI++)
    if(mask & (1u << sI))
    {
      const int   px = ox + SamplePattern<S>::xr(sI), py = oy + SamplePattern<S>::yr(sI);
      const Bary  b  = makeBary(edgeFloat(s, 1, px, py, small), edgeFloat(s, 2, px, py, small), s.rarea);
      const float zs = depthAt(s, b);
      if(!(zs < (dpx ? dpx[sI] : 1.0f)))
        mask &= ~(1u << sI);
    }
  return mask;
}

// Coverage mask of one pixel from the three edge functions e[] at the pixel's origin (fill-rule bias included) and the
// packed edge deltas pk[] = dx | -dy << 16 of a SMALL triangle: one IDP.2A per edge and sample evaluates
// e + dx * sy - dy * sx, the sign bits of the three are collected with one LOP3 + one funnel shift per sample.
template <int S>
__device__ __forceinline__ uint32_t sampleMaskSmall(const int e[3], const int pk[3])
{
  uint32_t outside = 0u;  // bit s: sample s fails an edge
#pragma unroll
  for(int sI = S - 1; sI >= 0; sI--)
  {
    const uint32_t sp = (uint32_t)SamplePattern<S>::y(sI) | ((uint32_t)SamplePattern<S>::x(sI) << 8);
    const int      e0 = dp2aS16U8(pk[0], sp, e[0]);
    const int      e1 = dp2aS16U8(pk[1], sp, e[1]);
    const int      e2 = dp2aS16U8(pk[2], sp, e[2]);
    outside           = __funnelshift_l((uint32_t)(e0 | e1 | e2), outside, 1);  // (outside << 1) | sign
  }
  return ~outside & ((1u << S) - 1u);
}

template <int S>
__device__ __forceinline__ uint32_t coverageMask(const TriSlot& s, int gx, int gy, const float* dpx)
{
  const bool small = (s.box >> 20) & 1u, zSafe = (s.box >> 19) & 1u;
  const int  ox = gx << 8, oy = gy << 8;
  uint32_t   mask = 0;
  if(small)
  {
    // |dx|, |dy| <= 2^14 fit a signed 16-bit half and the sample offsets an unsigned byte, so one IDP.2A per edge and
    // sample evaluates e + dx * sy - dy * sx
    int e[3], pk[3];
#pragma unroll
    for(int q = 0; q < 3; q++)
    {
      const int a = (q + 1) % 3, b = (q + 2) % 3;
      const int dx = s.x[b] - s.x[a], dy = s.y[b] - s.y[a];
      pk[q]       = (int)(((uint32_t)dx & 0xFFFFu) | ((uint32_t)(-dy) << 16));
      e[q]        = dx * (oy - s.y[a]) - dy * (ox - s.x[a]) - (int)((s.box >> (16 + q)) & 1u);
    }
#pragma unroll
    for(int sI = 0; sI < S; sI++)
    {
#if OIT_USE_DP2A
      const uint32_t sp = (uint32_t)SamplePattern<S>::y(sI) | ((uint32_t)SamplePattern<S>::x(sI) << 8);
      const int      e0 = dp2aS16U8(pk[0], sp, e[0]);
      const int      e1 = dp2aS16U8(pk[1], sp, e[1]);
      const int      e2 = dp2aS16U8(pk[2], sp, e[2]);
#else
      const int sx = SamplePattern<S>::x(sI), sy = SamplePattern<S>::y(sI);
      const int e0 = e[0] + (short)(pk[0] & 0xFFFF) * sy + (pk[0] >> 16) * sx;
      const int e1 = e[1] + (short)(pk[1] & 0xFFFF) * sy + (pk[1] >> 16) * sx;
      const int e2 = e[2] + (short)(pk[2] & 0xFFFF) * sy + (pk[2] >> 16) * sx;
#endif
      if((e0 | e1 | e2) >= 0)
        mask |= 1u << sI;
    }
  }
  else
    mask = coverageMaskLarge<S>(s, ox, oy);
  if(mask && (dpx != nullptr || !zSafe))
    mask = depthTestMask<S>(s, ox, oy, dpx, mask);
  return mask;
}

// Triangle set-up of one tile-list entry: orientation (area2 > 0), depth plane, fill-rule bias bits, the box of tile pixels
// that can hold a covered sample.  Returns the number of (triangle, pixel) items, padded to ITEMS_PER_THREAD.
// lo: smallest sample offset of the sample pattern in 1/256 px (samples sit in [lo, 256 - lo]).
__device__ __forceinline__ uint32_t setupSlot(TVert v0, TVert v1, TVert v2, uint32_t i0, uint32_t i1, uint32_t i2, uint32_t clipBits, int W, int H,
                                              int tileX0, int tileY0, int lo, TriSlot& s)
{
  const int hi    = 256 - lo;
  long long area2 = (long long)(v1.x - v0.x) * (v2.y - v0.y) - (long long)(v2.x - v0.x) * (v1.y - v0.y);
  if(area2 < 0)
  {
    const TVert tv = v1;
    v1             = v2;
    v2             = tv;
    if(clipBits == 0u)
    {
      const uint32_t ti = i1;
      i1                = i2;
      i2                = ti;
    }
    else
      clipBits |= SLOT_SWAPPED;  // vidx[0] stays the clip entry: shadeAt replays the exchange
    area2 = -area2;
  }
  const int minx = min(v0.x, min(v1.x, v2.x)), maxx = max(v0.x, max(v1.x, v2.x));
  const int miny = min(v0.y, min(v1.y, v2.y)), maxy = max(v0.y, max(v1.y, v2.y));
  const int px0 = max((minx - hi + 255) >> 8, tileX0), px1 = min((maxx - lo) >> 8, min(tileX0 + TILE_W, W) - 1);
  const int py0 = max((miny - hi + 255) >> 8, tileY0), py1 = min((maxy - lo) >> 8, min(tileY0 + TILE_H, H) - 1);
  // fill-rule bias of edge k (opposite vertex k, from vertex k+1 to vertex k+2): Vulkan / D3D top-left rule, y down, area2 > 0
  auto notTopLeft = [](int dx, int dy) { return ((dy == 0 && dx > 0) || dy < 0) ? 0u : 1u; };
  const uint32_t biasBits = notTopLeft(v2.x - v1.x, v2.y - v1.y) | (notTopLeft(v0.x - v2.x, v0.y - v2.y) << 1) | (notTopLeft(v1.x - v0.x, v1.y - v0.y) << 2);
  const bool     zSafe    = fmaxf(v0.z, fmaxf(v1.z, v2.z)) < 0.9999f;
  // extent <= 2^14 sub-pixels: |delta| <= 2^14, |sample - vertex| <= 2^14 + 2^12 inside the clipped box, so every
  // edge function and area fits comfortably in int32
  const bool small  = (maxx - minx) <= 16384 && (maxy - miny) <= 16384;
  uint32_t   box = 0, rcpW = 0, nItems = 0;
  if(px0 <= px1 && py0 <= py1)
  {
    const int bw = px1 - px0 + 1, bh = py1 - py0 + 1;
    // padded to a multiple of ITEMS_PER_THREAD so that the items of one thread always belong to one triangle
    nItems = (uint32_t)(bw * bh + ITEMS_PER_THREAD - 1) & ~(uint32_t)(ITEMS_PER_THREAD - 1);
    box    = (uint32_t)(px0 - tileX0) | ((uint32_t)(py0 - tileY0) << 4) | ((uint32_t)(bw - 1) << 8) | ((uint32_t)(bh - 1) << 12)
          | (biasBits << 16) | ((zSafe ? 1u : 0u) << 19) | ((small ? 1u : 0u) << 20) | clipBits;
    rcpW = (65535u + bw) / bw;
  }
  // s is the slot in shared memory: every field is written exactly once, from registers
  s.x[0] = v0.x; s.x[1] = v1.x; s.x[2] = v2.x;
  s.y[0] = v0.y; s.y[1] = v1.y; s.y[2] = v2.y;
  s.z0   = v0.z;
  s.dz1  = __fsub_rn(v1.z, v0.z);
  s.dz2  = __fsub_rn(v2.z, v0.z);
  s.iw[0] = v0.invw; s.iw[1] = v1.invw; s.iw[2] = v2.invw;
  s.vidx[0] = i0; s.vidx[1] = i1; s.vidx[2] = i2;
  s.rarea = __fdiv_rn(1.0f, __ll2float_rn(area2));
  s.box   = box;
  s.rcpW  = rcpW;
  return nItems;
}

// the technique a colour pass belongs to (the fused composite is specialised on it)
__host__ __device__ constexpr int passAlgorithm(int pass)
{
  return pass == PASS_SIMPLE ? OIT_SIMPLE
       : pass == PASS_LINKEDLIST ? OIT_LINKEDLIST
       : pass == PASS_LOOP_COLOR ? OIT_LOOP
       : pass == PASS_LOOP64 ? OIT_LOOP64
       : pass == PASS_SPINLOCK ? OIT_SPINLOCK
       : pass == PASS_INTERLOCK ? OIT_INTERLOCK
       : pass == PASS_WEIGHTED ? OIT_WEIGHTED
                               : -1;
}

// ---- phases shared by the batch kernels (oit_raster_ll.cu, oit_raster_q.cu) --------------------------------------------------
// A chunk stages CHUNK triangles (slots) whose candidates -- (triangle, pixel) pairs of the triangles' boxes inside the tile,
// row-major per triangle, padded to ITEMS_PER_THREAD -- form one item space (itemStart[slot] = first item of the slot).
//
// Coverage of the thread's ITEMS_PER_THREAD consecutive candidates k .. of one triangle: recs[j] = slot | lx << 8 | ly << 12 |
// mask << 16 (0 = not covered), and the covered pixel's bit `slot` is set in its per-batch triangle set (4 words per pixel).
template <int S, int CHUNK>
__device__ __forceinline__ void coverCandidates(const FrameParams& p, const TriSlot* slots, const uint32_t* itemStart, uint32_t k, uint32_t total,
                                                int tileX0, int tileY0, int yLocal0, uint32_t* setWords, uint32_t (&recs)[ITEMS_PER_THREAD])
{
#pragma unroll
  for(int j = 0; j < ITEMS_PER_THREAD; j++)
    recs[j] = 0u;
  if(k < total)
  {
    int slot = 0;
#pragma unroll
    for(int step = CHUNK / 2; step; step >>= 1)
      if(itemStart[slot + step] <= k)
        slot += step;
    const TriSlot& s      = slots[slot];
    const uint32_t box    = s.box;
    const uint32_t bw     = ((box >> 8) & 15u) + 1u, nPix = bw * (((box >> 12) & 15u) + 1u);
    const uint32_t local0 = k - itemStart[slot];
    uint32_t       row    = (local0 * s.rcpW) >> 16, col = local0 - row * bw;
    // all the masks first, the shared-memory atomics afterwards: nothing in between forces the slot to be read again
    uint32_t masks[ITEMS_PER_THREAD], pls[ITEMS_PER_THREAD];
    if(box & (1u << 20))
    {
      // extent <= 64 px: int32 edge functions, stepped from candidate to candidate (+1 px in x, or to the next box row)
      int e[3], pk[3], stepX[3], stepRow[3];  // stepX: one pixel to the right; stepRow: to the first pixel of the next box row
      {
        const int ox = (tileX0 + (int)(box & 15u) + (int)col) << 8, oy = (tileY0 + (int)((box >> 4) & 15u) + (int)row) << 8;
#pragma unroll
        for(int q = 0; q < 3; q++)
        {
          const int a = (q + 1) % 3, b = (q + 2) % 3;
          const int dx = s.x[b] - s.x[a], dy = s.y[b] - s.y[a];
          pk[q]       = (int)(((uint32_t)dx & 0xFFFFu) | ((uint32_t)(-dy) << 16));
          e[q]        = dx * (oy - s.y[a]) - dy * (ox - s.x[a]) - (int)((box >> (16 + q)) & 1u);
          stepX[q]    = -(dy << 8);
          stepRow[q]  = (dx << 8) + (dy << 8) * (int)(bw - 1u);
        }
      }
#pragma unroll
      for(int j = 0; j < ITEMS_PER_THREAD; j++)
      {
        pls[j]   = ((box & 15u) + col) | ((((box >> 4) & 15u) + row) << 4);
        masks[j] = 0u;
        if(local0 + j < nPix)
          masks[j] = sampleMaskSmall<S>(e, pk);
        if(j + 1 < ITEMS_PER_THREAD)
        {
          col++;
          const bool wrap = col == bw;
#pragma unroll
          for(int q = 0; q < 3; q++)
            e[q] += wrap ? stepRow[q] : stepX[q];
          if(wrap)
          {
            col = 0u;
            row++;
          }
        }
      }
    }
    else
    {
      // a triangle larger than 64 px: 64-bit edge functions, out of line (rare)
#pragma unroll
      for(int j = 0; j < ITEMS_PER_THREAD; j++)
      {
        pls[j]   = ((box & 15u) + col) | ((((box >> 4) & 15u) + row) << 4);
        masks[j] = 0u;
        if(local0 + j < nPix)
          masks[j] = coverageMaskLarge<S>(s, (tileX0 + (int)(pls[j] & 15u)) << 8, (tileY0 + (int)(pls[j] >> 4)) << 8);
        if(++col == bw)
        {
          col = 0u;
          row++;
        }
      }
    }
    if(p.depth != nullptr || !((box >> 19) & 1u))
    {
      // early per-sample depth test against the opaque pass (or a vertex depth close to the clear value): out of line
#pragma unroll
      for(int j = 0; j < ITEMS_PER_THREAD; j++)
        if(masks[j])
        {
          const int    lx = (int)(pls[j] & 15u), ly = (int)(pls[j] >> 4);
          const float* dpx = p.depth ? p.depth + ((size_t)(yLocal0 + ly) * p.W + tileX0 + lx) * S : nullptr;
          masks[j]         = depthTestMask<S>(s, (tileX0 + lx) << 8, (tileY0 + ly) << 8, dpx, masks[j]);
        }
    }
#pragma unroll
    for(int j = 0; j < ITEMS_PER_THREAD; j++)
      if(masks[j])
      {
        recs[j] = (uint32_t)slot | (pls[j] << 8) | (masks[j] << 16);
        atomicOr(&setWords[pls[j] * 4 + (slot >> 5)], 1u << (slot & 31));
      }
  }
}

// appends the thread's covered candidates to the batch's compact (unordered) list; *counter = length of the list
__device__ __forceinline__ void appendCovered(const uint32_t (&recs)[ITEMS_PER_THREAD], uint32_t* counter, uint32_t* list)
{
  const int lane = threadIdx.x & 31;
  uint32_t  cnt  = 0;
#pragma unroll
  for(int j = 0; j < ITEMS_PER_THREAD; j++)
    cnt += recs[j] ? 1u : 0u;
  const uint32_t incl  = warpInclusiveScan(cnt);
  const uint32_t wtot  = __shfl_sync(0xffffffffu, incl, 31);
  uint32_t       wbase = 0;
  if(lane == 31 && wtot)
    wbase = atomicAdd(counter, wtot);
  wbase        = __shfl_sync(0xffffffffu, wbase, 31);
  uint32_t pos = wbase + incl - cnt;
#pragma unroll
  for(int j = 0; j < ITEMS_PER_THREAD; j++)
    if(recs[j])
      list[pos++] = recs[j];
}

}  // namespace oit
