/*
 * oit_oracle.cpp -- CPU ORACLE (test infrastructure, see oit_oracle.h): a sequential restatement of the
 * reference's order-independent-transparency hot path, one function per reference shader / host stage.
 * Citations are file:line under /root/reference.  Pinned at image level against the reference's published
 * screenshot, UNPINNED below that (see header).
 *
 * Schedule: triangles in index order, and per triangle pixels in raster order: a legal schedule
 * of every technique and THE defined result for Loop32, ordered Interlock, Loop64 without tail blend and every
 * sorting technique that does not overflow.
 *
 * Arithmetic contract ("arith spec"): every float expression below is written as explicit
 * fmaf / single IEEE operations and the file is compiled with -ffp-contract=off, so that the CUDA path can be
 * bit-identical.  pow()-based sRGB conversions of the GLSL (shaderCommon.glsl:60-104) only ever touch 8-bit
 * values, so they are restated as exact tables: a 256-entry decode LUT and a 255-entry threshold table for the
 * encode; both are built in double precision from the sRGB formulas.
 */
#include "oit_oracle.h"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

enum { OIT_SIMPLE = 0, OIT_LINKEDLIST, OIT_LOOP, OIT_LOOP64, OIT_SPINLOCK, OIT_INTERLOCK, OIT_WEIGHTED, NUM_ALGORITHMS };
enum { AA_NONE = 0, AA_MSAA_4X, AA_SSAA_4X, AA_SUPER_4X, AA_MSAA_8X, AA_SSAA_8X, NUM_AATYPES };

/* ---- 8-bit sRGB tables (shaderCommon.glsl:60-104 restated exactly) ------------------------------------- */
double srgbToLinearD(double c) { return c <= 0.04045 ? c / 12.92 : std::pow((c + 0.055) / 1.055, 2.4); }

struct Tables
{
  float dec[256];  // sRGB8 code -> linear
  float thr[256];  // thr[k] = smallest linear value that encodes to >= k (k = 1..255); thr[0] = -inf
  float a255[256]; // v / 255.0f
  Tables()
  {
    for(int v = 0; v < 256; v++)
    {
      dec[v]  = (float)srgbToLinearD(v / 255.0);
      a255[v] = (float)v / 255.0f;
    }
    thr[0] = -std::numeric_limits<float>::infinity();
    for(int k = 1; k < 256; k++)
      thr[k] = (float)srgbToLinearD((k - 0.5) / 255.0);
  }
};
const Tables& T()
{
  static Tables t;
  return t;
}

inline uint32_t enc8(float c)
{
  const float* thr = T().thr;
  uint32_t     k   = 0;
  for(uint32_t step = 128; step; step >>= 1)
    if(c >= thr[k + step])
      k += step;
  return k;
}
inline float    clamp01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }
inline uint32_t unorm8(float a) { return (uint32_t)rintf(clamp01(a) * 255.0f); }
inline uint32_t fbits(float f)
{
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
inline float bitsf(uint32_t u)
{
  float f;
  memcpy(&f, &u, 4);
  return f;
}

/* packUnorm4x8(unPremultLinearToSRGB(c)): r in bits 0-7 ... a in 24-31 (oitSimple.frag.glsl:55,70) */
inline uint32_t packColor(const float c[4]) { return enc8(c[0]) | (enc8(c[1]) << 8) | (enc8(c[2]) << 16) | (unorm8(c[3]) << 24); }

/* unPremultSRGBToLinear(unpackUnorm4x8(p)) (shaderCommon.glsl:84-104) */
inline void unpackColor(uint32_t p, float c[4])
{
  const Tables& t = T();
  c[0]            = t.dec[p & 255];
  c[1]            = t.dec[(p >> 8) & 255];
  c[2]            = t.dec[(p >> 16) & 255];
  c[3]            = t.a255[p >> 24];
}
inline void premultiply(const float c[4], float o[4])
{
  o[0] = c[0] * c[3];
  o[1] = c[1] * c[3];
  o[2] = c[2] * c[3];
  o[3] = c[3];
}
/* doBlend (shaderCommon.glsl:108-112): color over base, premultiplied */
inline void doBlend(float color[4], const float base[4])
{
  const float t = 1.0f - color[3];
  color[0]      = fmaf(t, base[0], color[0]);
  color[1]      = fmaf(t, base[1], color[1]);
  color[2]      = fmaf(t, base[2], color[2]);
  color[3]      = fmaf(t, base[3], color[3]);
}
/* doBlendPacked (shaderCommon.glsl:117-124) */
inline void doBlendPacked(float color[4], uint32_t packed)
{
  float u[4], p[4];
  unpackColor(packed, u);
  premultiply(u, p);
  doBlend(color, p);
}

/* ---- half float (WBOIT targets are RGBA16F / R16F, oit.h:234-235) ---------------------------------------- */
uint16_t f2h(float f)
{
  uint32_t x    = fbits(f);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t ax   = x & 0x7fffffffu;
  if(ax >= 0x7f800000u)
    return (uint16_t)(sign | (ax > 0x7f800000u ? 0x7e00u : 0x7c00u));
  if(ax >= 0x477ff000u)  // rounds to >= 65520 -> inf
    return (uint16_t)(sign | 0x7c00u);
  if(ax < 0x33000001u)  // < 2^-25 (or exactly 2^-25 which ties to even 0)
    return (uint16_t)sign;
  int32_t  e = (int32_t)(ax >> 23) - 127;
  uint32_t m = (ax & 0x7fffffu) | 0x800000u;
  int      shift;
  uint32_t base;
  if(e < -14)
  {
    shift = 13 + (-14 - e);
    base  = 0;
  }
  else
  {
    shift = 13;
    base  = (uint32_t)(e + 15) << 10;
    m &= 0x7fffffu;
  }
  uint32_t q    = m >> shift;
  uint32_t rem  = m & ((1u << shift) - 1u);
  uint32_t half = 1u << (shift - 1);
  if(rem > half || (rem == half && (q & 1u)))
    q++;
  return (uint16_t)(sign | (base + q));
}
float h2f(uint16_t h)
{
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t e    = (h >> 10) & 31u;
  uint32_t m    = h & 1023u;
  if(e == 31)
    return bitsf(sign | 0x7f800000u | (m << 13));
  if(e == 0)
  {
    if(m == 0)
      return bitsf(sign);
    float v = (float)m * (1.0f / 16777216.0f);  // m * 2^-24
    return sign ? -v : v;
  }
  return bitsf(sign | ((e + 112u) << 23) | (m << 13));
}

/* ---- ROP: B8G8R8A8_SRGB target (oit.cpp:58), blend modes main.cpp:540-592 ---------------------------------- */
inline void decodeDst(uint32_t d, float c[4])
{
  const Tables& t = T();
  c[0]            = t.dec[(d >> 16) & 255];  // R
  c[1]            = t.dec[(d >> 8) & 255];   // G
  c[2]            = t.dec[d & 255];          // B
  c[3]            = t.a255[d >> 24];
}
inline uint32_t encodeDst(const float c[4]) { return enc8(c[2]) | (enc8(c[1]) << 8) | (enc8(c[0]) << 16) | (unorm8(c[3]) << 24); }
/* BlendMode::PREMULTIPLIED: dst = src + (1-src.a)*dst for rgb and a (main.cpp:548-558) */
inline uint32_t ropPremult(uint32_t dst, const float src[4])
{
  float d[4], r[4];
  decodeDst(dst, d);
  const float t = 1.0f - src[3];
  for(int i = 0; i < 4; i++)
    r[i] = fmaf(t, d[i], src[i]);
  return encodeDst(r);
}
/* BlendMode::WEIGHTED_COMPOSITE: (1-srcA)*src + srcA*dst (main.cpp:576-588) */
inline uint32_t ropWeightedComposite(uint32_t dst, const float src[4])
{
  float d[4], r[4];
  decodeDst(dst, d);
  const float om = 1.0f - src[3];
  for(int i = 0; i < 4; i++)
    r[i] = fmaf(src[3], d[i], om * src[i]);
  return encodeDst(r);
}

/* ---- geometry ---------------------------------------------------------------------------------------- */
struct TVert
{
  int32_t x, y;  // framebuffer position, 8 sub-pixel bits
  float   z, invw, viewz;
  bool    valid;
};
const int SPOS1[1][2] = {{128, 128}};
const int SPOS4[4][2] = {{96, 32}, {224, 96}, {32, 160}, {160, 224}};
const int SPOS8[8][2] = {{144, 80}, {112, 176}, {208, 144}, {80, 48}, {48, 208}, {16, 112}, {176, 240}, {240, 16}};

struct Frag
{
  uint32_t x, y, sampleID, mask;  // mask = post-depth coverage (gl_SampleMaskIn[0])
  float    rgba[4];               // shading(): unpremultiplied linear colour
  float    z;                     // gl_FragCoord.z
  float    viewz;                 // IN.depth
};

}  // namespace

struct OracleCtx
{
  OracleConfig cfg;
  int          msaa = 1, supersample = 1;
  bool         sampleShading = false, coverage = false;
  uint32_t     W = 0, H = 0;  // render-target size (after supersample)
  uint32_t     layers = 1;    // A-buffer / aux layers
  uint32_t     L      = 8;
  uint32_t     capacity = 0;  // linked list: scene.linkedListAllocatedPerElement (oit.cpp:125,151)
  int          threads  = 1;

  std::vector<float>    verts;
  std::vector<uint32_t> idx;
  uint32_t              idxPerObj = 0;
  OracleSceneData       ubo;
  std::vector<TVert>    tv;

  std::vector<uint32_t> abuf, aux, spin, adepth;
  std::atomic<uint32_t> counter{0};
  uint32_t              counterOut = 0;
  std::vector<uint32_t> color;
  std::vector<float>    depth;
  std::vector<uint16_t> wacc, wrev;
  std::vector<uint32_t> fin;
  OracleStats           stats;
  const int (*spos)[2] = SPOS1;
};

namespace {

/* the vertex stage after the matrix product: perspective divide, viewport transform (main.cpp:507-510), guard band,
   snapping to 1/256 px */
TVert finishVertex(const float clip[4], float viewz, float hw, float hh)
{
  TVert t;
  t.viewz = viewz;
  t.valid = clip[3] > 0.f && clip[3] < std::numeric_limits<float>::infinity();
  t.x = t.y = 0;
  t.z = t.invw = 0.f;
  if(t.valid)
  {
    t.invw         = 1.0f / clip[3];
    const float nx = clip[0] * t.invw, ny = clip[1] * t.invw;
    t.z            = clip[2] * t.invw;
    const float xs = fmaf(nx, hw, hw), ys = fmaf(ny, hh, hh);
    if(!(fabsf(xs) < 2097152.f) || !(fabsf(ys) < 2097152.f) || !(t.z >= 0.f) || !(t.z <= 1.f))
      t.valid = false;  // outside the guard band / depth clip volume
    else
    {
      t.x = (int32_t)rintf(xs * 256.0f);
      t.y = (int32_t)rintf(ys * 256.0f);
    }
  }
  if(!t.valid)
  {
    t.x = t.y = 0;
    t.z = t.invw = 0.f;
  }
  return t;
}

/* ---- vertex stage: object.vert.glsl:32-38 + viewport transform (main.cpp:507-510) ------------------------- */
void transformVertices(OracleCtx* c)
{
  const size_t n = c->verts.size() / 10;
  c->tv.resize(n);
  const float* M  = c->ubo.projViewMatrix;
  const float* V  = c->ubo.viewMatrix;
  const float  hw = 0.5f * (float)c->W, hh = 0.5f * (float)c->H;
#pragma omp parallel for num_threads(c->threads) if(c->threads > 1)
  for(long i = 0; i < (long)n; i++)
  {
    const float* p = &c->verts[i * 10];
    float        clip[4];
    for(int r = 0; r < 4; r++)
      clip[r] = fmaf(M[0 + r], p[0], fmaf(M[4 + r], p[1], fmaf(M[8 + r], p[2], M[12 + r])));
    const float viewz = fmaf(V[0 + 2], p[0], fmaf(V[4 + 2], p[1], fmaf(V[8 + 2], p[2], V[12 + 2])));
    const TVert t     = finishVertex(clip, viewz, hw, hh);
    c->tv[i] = t;
  }
}

/* shading() + goochLighting() (shaderCommon.glsl:36-56) */
inline void shade(const OracleCtx* c, const float n[3], const float col[4], float out[4])
{
  const float LX = -0.40824829046386301637f, LY = 0.81649658092772603273f, LZ = 0.40824829046386301637f;
  const float len2 = fmaf(n[2], n[2], fmaf(n[1], n[1], n[0] * n[0]));
  float       nx = 0.f, ny = 0.f, nz = 0.f;
  if(len2 > 0.f)
  {
    const float inv = 1.0f / sqrtf(len2);
    nx              = n[0] * inv;
    ny              = n[1] * inv;
    nz              = n[2] * inv;
  }
  const float d      = fmaf(nz, LZ, fmaf(ny, LY, nx * LX));
  const float warmth = fmaf(d, 0.5f, 0.5f);
  const float om     = 1.0f - warmth;
  out[0]             = col[0] * fmaf(0.0f, om, warmth);
  out[1]             = col[1] * fmaf(0.25f, om, warmth);
  out[2]             = col[2] * fmaf(0.75f, om, warmth);
  out[3]             = clamp01(fmaf(col[3], c->ubo.alphaWidth, c->ubo.alphaMin));
}

struct ThreadStats
{
  uint64_t fragments = 0, stored = 0, tail = 0, opaque = 0, rejected = 0;
};

inline size_t auxIndex(const OracleCtx* c, const Frag& f) { return ((size_t)f.sampleID * c->H + f.y) * c->W + f.x; }

/* ROP write of a colour-pass / composite output onto the samples in `mask` of pixel (x,y) */
inline void ropToSamples(OracleCtx* c, uint32_t x, uint32_t y, uint32_t mask, const float src[4])
{
  if(src[0] == 0.f && src[1] == 0.f && src[2] == 0.f && src[3] == 0.f)
    return;  // identity blend (tests check encode(decode(v)) == v)
  uint32_t* px = &c->color[((size_t)y * c->W + x) * c->msaa];
  for(int s = 0; s < c->msaa; s++)
    if(mask & (1u << s))
      px[s] = ropPremult(px[s], src);
}

/* ---- colour passes -------------------------------------------------------------------------------------- */

/* oitSimple.frag.glsl:50-91 */
void colorSimple(OracleCtx* c, const Frag& f, float out[4], ThreadStats& st)
{
  const size_t   viewSize = (size_t)c->W * c->H;
  const size_t   listPos  = viewSize * c->L * f.sampleID + (size_t)f.y * c->W + f.x;
  const uint32_t stride   = c->coverage ? 4 : 2;
  const uint32_t old      = c->aux[auxIndex(c, f)]++;
  if(old < c->L)
  {
    uint32_t* e = &c->abuf[(listPos + (size_t)old * viewSize) * stride];
    e[0]        = packColor(f.rgba);
    e[1]        = fbits(f.z);
    if(c->coverage)
    {
      e[2] = f.mask;
      e[3] = 0;
    }
    out[0] = out[1] = out[2] = out[3] = 0.f;
    st.stored++;
  }
  else if(c->cfg.tailBlend)
  {
    premultiply(f.rgba, out);
    st.tail++;
  }
  else
    out[0] = out[1] = out[2] = out[3] = 0.f;
}

/* oitLinkedList.frag.glsl:51-85 */
void colorLinkedList(OracleCtx* c, const Frag& f, float out[4], ThreadStats& st)
{
  const uint32_t newOffset = c->counter.fetch_add(1u, std::memory_order_relaxed) + 1u;
  if(newOffset >= c->capacity)
  {
    if(c->cfg.tailBlend)
    {
      premultiply(f.rgba, out);
      st.tail++;
    }
    else
      out[0] = out[1] = out[2] = out[3] = 0.f;
    return;
  }
  uint32_t&      head = c->aux[auxIndex(c, f)];
  const uint32_t old  = head;
  head                = newOffset;
  uint32_t* e         = &c->abuf[(size_t)newOffset * 4];
  e[0]                = packColor(f.rgba);
  e[1]                = fbits(f.z);
  e[2]                = c->coverage ? f.mask : 0u;
  e[3]                = old;
  out[0] = out[1] = out[2] = out[3] = 0.f;
  st.stored++;
}

/* oitLoop.frag.glsl:57-100 (depth pass) */
void depthLoop(OracleCtx* c, const Frag& f)
{
  const size_t viewSize = (size_t)c->W * c->H;
  const size_t listPos  = viewSize * c->L * 2 * f.sampleID + (size_t)f.y * c->W + f.x;
  uint32_t     zcur     = fbits(f.z);
  uint32_t     i        = 0;
  uint32_t     pretest  = c->abuf[listPos + (size_t)(c->L - 1) * viewSize];
  if(zcur > pretest)
    return;
  pretest = c->abuf[listPos + (size_t)(c->L / 2) * viewSize];
  if(zcur > pretest)
    i = c->L / 2;
  for(; i < c->L; i++)
  {
    uint32_t&      slot  = c->abuf[listPos + (size_t)i * viewSize];
    const uint32_t ztest = slot;
    slot                 = std::min(ztest, zcur);
    if(ztest == 0xFFFFFFFFu || ztest == zcur)
      break;
    zcur = std::max(ztest, zcur);
  }
}
/* oitLoop.frag.glsl:120-173 (colour pass) */
void colorLoop(OracleCtx* c, const Frag& f, float out[4], ThreadStats& st)
{
  const size_t   viewSize = (size_t)c->W * c->H;
  const size_t   listPos  = viewSize * c->L * 2 * f.sampleID + (size_t)f.y * c->W + f.x;
  const uint32_t zcur     = fbits(f.z);
  out[0] = out[1] = out[2] = out[3] = 0.f;
  if(c->abuf[listPos + (size_t)(c->L - 1) * viewSize] < zcur)
  {
    if(c->cfg.tailBlend)
    {
      premultiply(f.rgba, out);
      st.tail++;
    }
    return;
  }
  int start = 0, end = (int)c->L - 1;
  while(start < end)
  {
    const int      mid   = (start + end) / 2;
    const uint32_t ztest = c->abuf[listPos + (size_t)mid * viewSize];
    if(ztest < zcur)
      start = mid + 1;
    else
      end = mid;
  }
  c->abuf[listPos + (size_t)(c->L + start) * viewSize] = packColor(f.rgba);
  st.stored++;
}

/* oitLoop64.frag.glsl:65-141 */
void colorLoop64(OracleCtx* c, const Frag& f, float out[4], ThreadStats& st)
{
  const size_t viewSize = (size_t)c->W * c->H;
  const size_t listPos  = viewSize * c->L * f.sampleID + (size_t)f.y * c->W + f.x;
  uint64_t*    ab       = reinterpret_cast<uint64_t*>(c->abuf.data());
  uint64_t     zcur     = ((uint64_t)fbits(f.z) << 32) | packColor(f.rgba);
  uint32_t     i        = 0;
  bool         canInsert = true;
  uint64_t     pretest   = ab[listPos + (size_t)(c->L - 1) * viewSize];
  if(zcur > pretest)
    canInsert = false;
  else
  {
    pretest = ab[listPos + (size_t)(c->L / 2) * viewSize];
    if(zcur > pretest)
      i = c->L / 2;
  }
  bool evict = true;
  if(canInsert)
  {
    for(; i < c->L; i++)
    {
      uint64_t&      slot  = ab[listPos + (size_t)i * viewSize];
      const uint64_t ztest = slot;
      slot                 = std::min(ztest, zcur);
      if(ztest == ~0ull)
      {
        evict = false;
        break;
      }
      zcur = (ztest > zcur) ? ztest : zcur;
    }
  }
  out[0] = out[1] = out[2] = out[3] = 0.f;
  if(!evict)
  {
    st.stored++;
    return;
  }
  if(c->cfg.tailBlend)
  {
    float u[4];
    unpackColor((uint32_t)(zcur & 0xFFFFFFFFu), u);
    premultiply(u, out);
    st.tail++;
  }
}

/* oitSpinlock.frag.glsl:49-129 and oitInterlock.frag.glsl:90-152: under the sequential schedule the lock is
   always free and the early-out read of imgDepth sees the same value inside or outside the critical section. */
void colorLock(OracleCtx* c, const Frag& f, float out[4], ThreadStats& st)
{
  const size_t   viewSize = (size_t)c->W * c->H;
  const size_t   listPos  = viewSize * c->L * f.sampleID + (size_t)f.y * c->W + f.x;
  const uint32_t stride   = c->coverage ? 4 : 2;
  const size_t   ai       = auxIndex(c, f);
  const uint32_t zbits    = fbits(f.z);
  float          color[4] = {f.rgba[0], f.rgba[1], f.rgba[2], f.rgba[3]};
  bool           stored   = false;
  if(zbits <= c->adepth[ai] && f.mask != 0)
  {
    const uint32_t oldCounter = c->aux[ai];
    c->aux[ai]                = oldCounter + 1;
    uint32_t sv[4]            = {packColor(f.rgba), zbits, c->coverage ? f.mask : 0u, 0u};
    if(oldCounter < c->L)
    {
      memcpy(&c->abuf[(listPos + (size_t)oldCounter * viewSize) * stride], sv, stride * 4);
      color[0] = color[1] = color[2] = color[3] = 0.f;
      stored                                    = true;
    }
    else
    {
      uint32_t furthest = 0, maxDepth = 0;
      for(uint32_t i = 0; i < c->L; i++)
      {
        const uint32_t testDepth = c->abuf[(listPos + (size_t)i * viewSize) * stride + 1];
        if(testDepth > maxDepth)
        {
          maxDepth = testDepth;
          furthest = i;
        }
      }
      if(maxDepth > zbits)
      {
        uint32_t* e = &c->abuf[(listPos + (size_t)furthest * viewSize) * stride];
        unpackColor(e[0], color);
        memcpy(e, sv, stride * 4);
        c->adepth[ai] = maxDepth;
        stored        = true;  // this fragment went in; the evicted one is what tail blends
      }
    }
  }
  if(stored)
    st.stored++;
  if(c->cfg.tailBlend)
  {
    premultiply(color, out);
    if(out[3] != 0.f || out[0] != 0.f || out[1] != 0.f || out[2] != 0.f)
      st.tail++;
  }
  else
    out[0] = out[1] = out[2] = out[3] = 0.f;  // outColor is never written (oitSpinlock.frag.glsl:126-128): defined as 0
}

/* oitWeighted.frag.glsl:53-79 + BlendMode::WEIGHTED_COLOR (main.cpp:559-575) */
void colorWeighted(OracleCtx* c, const Frag& f, ThreadStats& st)
{
  float col[4];
  premultiply(f.rgba, col);
  const float depthZ     = -f.viewz * 10.0f;
  const float x          = depthZ / 200.0f;
  const float x2         = x * x;
  const float x4         = x2 * x2;
  float       distWeight = 0.03f / (1e-5f + x4);
  distWeight             = distWeight < 1e-2f ? 1e-2f : (distWeight > 3e3f ? 3e3f : distWeight);
  const float mx         = std::max(std::max(col[0], col[1]), std::max(col[2], col[3]));
  float       aw         = std::min(1.0f, fmaf(mx, 40.0f, 0.01f));
  aw                     = aw * aw;
  const float weight     = aw * distWeight;
  const float om         = 1.0f - col[3];
  uint16_t*   acc        = &c->wacc[((size_t)f.y * c->W + f.x) * c->msaa * 4];
  uint16_t*   rev        = &c->wrev[((size_t)f.y * c->W + f.x) * c->msaa];
  for(int s = 0; s < c->msaa; s++)
    if(f.mask & (1u << s))
    {
      for(int k = 0; k < 4; k++)
        acc[s * 4 + k] = f2h(h2f(acc[s * 4 + k]) + col[k] * weight);
      rev[s] = f2h(h2f(rev[s]) * om);
    }
  st.stored++;
}

/* one colour-pass invocation + its ROP write */
inline void invoke(OracleCtx* c, int pass, const Frag& f, ThreadStats& st, float* outOpt = nullptr)
{
  float out[4] = {0, 0, 0, 0};
  if(pass == 0)
  {
    depthLoop(c, f);
    if(outOpt)
      memcpy(outOpt, out, 16);
    return;
  }
  st.fragments++;
  switch(c->cfg.algorithm)
  {
    case OIT_SIMPLE: colorSimple(c, f, out, st); break;
    case OIT_LINKEDLIST: colorLinkedList(c, f, out, st); break;
    case OIT_LOOP: colorLoop(c, f, out, st); break;
    case OIT_LOOP64: colorLoop64(c, f, out, st); break;
    case OIT_SPINLOCK:
    case OIT_INTERLOCK: colorLock(c, f, out, st); break;
    case OIT_WEIGHTED: colorWeighted(c, f, st); break;
  }
  if(outOpt)
    memcpy(outOpt, out, 16);
  if(c->cfg.algorithm != OIT_WEIGHTED)
    ropToSamples(c, f.x, f.y, f.mask, out);
}

/* ---- rasteriser: fixed-function rules of SURVEY 8a row R --------------------------------------------------- */
struct Tri
{
  int64_t      x[3], y[3];
  float        z[3], iw[3], vz[3];
  const float* a[3];  // -> vertex (pos3, normal3, colour4)
  int64_t      area2;
  int          bias[3];
};

inline bool topLeft(int64_t dx, int64_t dy) { return (dy == 0 && dx > 0) || dy < 0; }

/* triangle set-up from three post-projection vertices and their attribute records (pos3, normal3, colour4) */
bool setupTriFrom(const TVert* const v[3], const float* const attr[3], bool cullBack, Tri& t)
{
  int64_t area2 = ((int64_t)v[1]->x - v[0]->x) * ((int64_t)v[2]->y - v[0]->y) - ((int64_t)v[2]->x - v[0]->x) * ((int64_t)v[1]->y - v[0]->y);
  if(area2 == 0)
    return false;
  /* Vulkan: a = -1/2 sum(x_i*y_i+1 - x_i+1*y_i); positive = front for COUNTER_CLOCKWISE => front iff area2 < 0 */
  if(cullBack && area2 > 0)
    return false;
  int o[3] = {0, 1, 2};
  if(area2 < 0)
  {
    std::swap(o[1], o[2]);
    area2 = -area2;
  }
  for(int k = 0; k < 3; k++)
  {
    const TVert& tv = *v[o[k]];
    t.x[k]          = tv.x;
    t.y[k]          = tv.y;
    t.z[k]          = tv.z;
    t.iw[k]         = tv.invw;
    t.vz[k]         = tv.viewz;
    t.a[k]          = attr[o[k]];
  }
  t.area2 = area2;
  /* edge k is opposite vertex k: from vertex k+1 to vertex k+2 */
  for(int k = 0; k < 3; k++)
  {
    const int a = (k + 1) % 3, b = (k + 2) % 3;
    t.bias[k]   = topLeft(t.x[b] - t.x[a], t.y[b] - t.y[a]) ? 0 : -1;
  }
  return true;
}

/* Near-plane clipping (the fixed-function clipper, 0 <= z_clip) of a triangle with a vertex behind the
   near plane.  Rules (the CUDA side states the same ones in csrc/oit_clip.cuh):
    - inside iff z_clip >= 0, clip coordinates from the vertex stage's fma chain;
    - a new vertex lies on an edge from an INSIDE vertex P to an OUTSIDE vertex Q, always computed in that direction:
      t = zP / (zP - zQ); x, y, w, view-z and the attributes = fma(t, Q - P, P); z_clip := 0;
    - one vertex outside (k; a = k+1, b = k+2): A' on a->k, B' on b->k, pieces (A', a, b) and (A', b, B');
      two outside (inside a; b = a+1, c = a+2): P on a->b, Q on a->c, piece (a, P, Q);
    - any vertex not representable afterwards (w <= 0, guard band, z outside [0,1]) => the triangle stays rejected. */
struct ClipPiece
{
  TVert v[3];
  float attr[3][10];
};
int clipTriangleNear(const OracleCtx* c, const uint32_t ix[3], ClipPiece out[2])
{
  const float* M  = c->ubo.projViewMatrix;
  const float* V  = c->ubo.viewMatrix;
  const float  hw = 0.5f * (float)c->W, hh = 0.5f * (float)c->H;
  float        clip[3][4], vz[3];
  int          nIn = 0, firstOut = -1, firstIn = -1;
  for(int k = 0; k < 3; k++)
  {
    const float* p = &c->verts[(size_t)ix[k] * 10];
    for(int r = 0; r < 4; r++)
      clip[k][r] = fmaf(M[0 + r], p[0], fmaf(M[4 + r], p[1], fmaf(M[8 + r], p[2], M[12 + r])));
    vz[k] = fmaf(V[0 + 2], p[0], fmaf(V[4 + 2], p[1], fmaf(V[8 + 2], p[2], V[12 + 2])));
    if(clip[k][2] >= 0.f)
    {
      nIn++;
      if(firstIn < 0)
        firstIn = k;
    }
    else if(firstOut < 0)
      firstOut = k;
  }
  if(nIn == 0 || nIn == 3)
    return 0;
  struct CV
  {
    TVert v;
    float attr[10];
  };
  auto original = [&](int k) {
    CV cv;
    cv.v = finishVertex(clip[k], vz[k], hw, hh);
    memcpy(cv.attr, &c->verts[(size_t)ix[k] * 10], sizeof(cv.attr));
    return cv;
  };
  auto cut = [&](int in, int outV) {
    const float zP = clip[in][2], zQ = clip[outV][2];
    const float t  = zP / (zP - zQ);
    float       c4[4];
    c4[0] = fmaf(t, clip[outV][0] - clip[in][0], clip[in][0]);
    c4[1] = fmaf(t, clip[outV][1] - clip[in][1], clip[in][1]);
    c4[2] = 0.f;
    c4[3] = fmaf(t, clip[outV][3] - clip[in][3], clip[in][3]);
    CV cv;
    cv.v             = finishVertex(c4, fmaf(t, vz[outV] - vz[in], vz[in]), hw, hh);
    const float* aP = &c->verts[(size_t)ix[in] * 10];
    const float* aQ = &c->verts[(size_t)ix[outV] * 10];
    for(int k = 0; k < 10; k++)
      cv.attr[k] = fmaf(t, aQ[k] - aP[k], aP[k]);
    return cv;
  };
  int  count = 0;
  CV   pv[2][3];
  if(nIn == 2)
  {
    const int k = firstOut, a = (k + 1) % 3, b = (k + 2) % 3;
    const CV  A = cut(a, k), B = cut(b, k), va = original(a), vb = original(b);
    pv[0][0] = A;
    pv[0][1] = va;
    pv[0][2] = vb;
    pv[1][0] = A;
    pv[1][1] = vb;
    pv[1][2] = B;
    count    = 2;
  }
  else
  {
    const int a = firstIn, b = (a + 1) % 3, cc = (a + 2) % 3;
    pv[0][0] = original(a);
    pv[0][1] = cut(a, b);
    pv[0][2] = cut(a, cc);
    count    = 1;
  }
  for(int s = 0; s < count; s++)
    for(int k = 0; k < 3; k++)
    {
      if(!pv[s][k].v.valid)
        return 0;
      out[s].v[k] = pv[s][k].v;
      memcpy(out[s].attr[k], pv[s][k].attr, sizeof(pv[s][k].attr));
    }
  return count;
}

inline int64_t edgeFn(const Tri& t, int k, int64_t px, int64_t py)
{
  const int a = (k + 1) % 3, b = (k + 2) % 3;
  return (t.x[b] - t.x[a]) * (py - t.y[a]) - (t.y[b] - t.y[a]) * (px - t.x[a]);
}
struct Bary
{
  float l0, l1, l2;
};
inline Bary baryAt(const Tri& t, int64_t px, int64_t py)
{
  /* screen-space barycentrics: edge function times the reciprocal of the doubled area (one IEEE division per triangle) */
  const float ra = 1.0f / (float)t.area2;
  Bary        b;
  b.l1 = (float)edgeFn(t, 1, px, py) * ra;
  b.l2 = (float)edgeFn(t, 2, px, py) * ra;
  b.l0 = (1.0f - b.l1) - b.l2;
  return b;
}
inline float depthAt(const Tri& t, const Bary& b) { return clamp01(fmaf(b.l2, t.z[2] - t.z[0], fmaf(b.l1, t.z[1] - t.z[0], t.z[0]))); }

/* perspective-correct varyings (Interpolants, shaderCommon.glsl:25-31): normal, colour, view z */
inline void varyingsAt(const Tri& t, const Bary& b, float n[3], float col[4], float& viewz)
{
  const float q0 = b.l0 * t.iw[0], q1 = b.l1 * t.iw[1], q2 = b.l2 * t.iw[2];
  const float rden = 1.0f / ((q0 + q1) + q2);
  for(int k = 0; k < 3; k++)
    n[k] = fmaf(q2, t.a[2][3 + k], fmaf(q1, t.a[1][3 + k], q0 * t.a[0][3 + k])) * rden;
  for(int k = 0; k < 4; k++)
    col[k] = fmaf(q2, t.a[2][6 + k], fmaf(q1, t.a[1][6 + k], q0 * t.a[0][6 + k])) * rden;
  viewz = fmaf(q2, t.vz[2], fmaf(q1, t.vz[1], q0 * t.vz[0])) * rden;
}

void rasterTri(OracleCtx* c, const Tri& t, int mode, int rowBegin, int rowEnd, ThreadStats& st)
{
  const int64_t minx = std::min(t.x[0], std::min(t.x[1], t.x[2])), maxx = std::max(t.x[0], std::max(t.x[1], t.x[2]));
  const int64_t miny = std::min(t.y[0], std::min(t.y[1], t.y[2])), maxy = std::max(t.y[0], std::max(t.y[1], t.y[2]));
  int           px0 = (int)std::max<int64_t>(minx >> 8, 0), px1 = (int)std::min<int64_t>(maxx >> 8, (int64_t)c->W - 1);
  int           py0 = (int)std::max<int64_t>(miny >> 8, rowBegin), py1 = (int)std::min<int64_t>(maxy >> 8, (int64_t)rowEnd - 1);
  const int     S        = c->msaa;
  const bool    weighted = c->cfg.algorithm == OIT_WEIGHTED;
  const bool    perSample = (mode != 0) && c->sampleShading && !weighted;  // oitColorDepthDefines.glsl:49-57
  for(int py = py0; py <= py1; py++)
    for(int px = px0; px <= px1; px++)
    {
      float*   dpx  = &c->depth[((size_t)py * c->W + px) * S];
      uint32_t mask = 0;
      float    zs[8];
      Bary     bs[8];
      for(int s = 0; s < S; s++)
      {
        const int64_t sx = (int64_t)px * 256 + c->spos[s][0], sy = (int64_t)py * 256 + c->spos[s][1];
        if(edgeFn(t, 0, sx, sy) + t.bias[0] < 0 || edgeFn(t, 1, sx, sy) + t.bias[1] < 0 || edgeFn(t, 2, sx, sy) + t.bias[2] < 0)
          continue;
        bs[s] = baryAt(t, sx, sy);
        zs[s] = depthAt(t, bs[s]);
        if(zs[s] < dpx[s])  // VK_COMPARE_OP_LESS (main.cpp:530-532)
          mask |= 1u << s;
      }
      if(!mask)
        continue;
      if(mode == 0)
      {
        /* opaque.frag.glsl:30-34; BlendMode::NONE + depth write (main.cpp:541-546); per-pixel shading */
        const Bary bc = baryAt(t, (int64_t)px * 256 + 128, (int64_t)py * 256 + 128);
        float      n[3], col[4], vz, g[4];
        varyingsAt(t, bc, n, col, vz);
        shade(c, n, col, g);
        g[3]               = 1.0f;
        const uint32_t enc = encodeDst(g);
        uint32_t*      cpx = &c->color[((size_t)py * c->W + px) * S];
        for(int s = 0; s < S; s++)
          if(mask & (1u << s))
          {
            cpx[s] = enc;
            dpx[s] = zs[s];
          }
        st.opaque++;
        continue;
      }
      Frag f;
      f.x = px;
      f.y = py;
      if(perSample)
      {
        for(int s = 0; s < S; s++)
          if(mask & (1u << s))
          {
            float n[3], col[4];
            varyingsAt(t, bs[s], n, col, f.viewz);
            shade(c, n, col, f.rgba);
            f.sampleID = s;
            f.mask     = 1u << s;
            f.z        = zs[s];
            invoke(c, mode == 1 ? 0 : 1, f, st);
          }
      }
      else
      {
        const Bary bc = (S == 1) ? bs[0] : baryAt(t, (int64_t)px * 256 + 128, (int64_t)py * 256 + 128);
        float      n[3], col[4];
        varyingsAt(t, bc, n, col, f.viewz);
        shade(c, n, col, f.rgba);
        f.sampleID = 0;
        f.mask     = mask;
        f.z        = (S == 1) ? zs[0] : depthAt(t, bc);
        invoke(c, mode == 1 ? 0 : 1, f, st);
      }
    }
}

void drawRange(OracleCtx* c, uint32_t firstIndex, uint32_t indexCount, int mode)
{
  const int               nt = std::max(1, c->threads);
  std::vector<ThreadStats> sts(nt);
#pragma omp parallel num_threads(nt) if(nt > 1)
  {
#ifdef _OPENMP
    const int tid = omp_get_thread_num();
#else
    const int tid = 0;
#endif
    /* band-parallel: each thread owns a contiguous band of rows and walks ALL triangles in order, so the
       per-pixel order is primitive order in every band */
    const int    rows0 = (int)((uint64_t)c->H * tid / nt), rows1 = (int)((uint64_t)c->H * (tid + 1) / nt);
    ThreadStats& st = sts[tid];
    Tri          t;
    ClipPiece    pieces[2];
    for(uint32_t i = 0; i + 2 < indexCount; i += 3)
    {
      const uint32_t* ix = &c->idx[firstIndex + i];
      const TVert* const v[3] = {&c->tv[ix[0]], &c->tv[ix[1]], &c->tv[ix[2]]};
      if(v[0]->valid && v[1]->valid && v[2]->valid)
      {
        const float* const attr[3] = {&c->verts[(size_t)ix[0] * 10], &c->verts[(size_t)ix[1] * 10], &c->verts[(size_t)ix[2] * 10]};
        if(setupTriFrom(v, attr, mode == 0, t))
          rasterTri(c, t, mode, rows0, rows1, st);
        continue;
      }
      /* a vertex has no post-projection position: behind the near plane => clip; anything else => rejected */
      const int n = clipTriangleNear(c, ix, pieces);
      if(n == 0 && mode != 1)  /* counted once per draw: Loop32's depth pre-pass walks the same triangles again */
        st.rejected++;
      for(int s = 0; s < n; s++)
      {
        const TVert* const pvv[3]  = {&pieces[s].v[0], &pieces[s].v[1], &pieces[s].v[2]};
        const float* const pat[3] = {pieces[s].attr[0], pieces[s].attr[1], pieces[s].attr[2]};
        if(setupTriFrom(pvv, pat, mode == 0, t))
          rasterTri(c, t, mode, rows0, rows1, st);
      }
    }
  }
  for(int i = 0; i < nt; i++)
  {
    c->stats.fragments += sts[i].fragments;
    c->stats.fragmentsStored += sts[i].stored;
    c->stats.fragmentsTail += sts[i].tail;
    c->stats.opaqueFragments += sts[i].opaque;
  }
  c->stats.trianglesRejected += sts[0].rejected;
}

/* ---- composites --------------------------------------------------------------------------------------- */
struct Elem
{
  uint32_t c, d, m;
};
/* bubbleSort (oitCompositeDefines.glsl:51-89): swaps on >= of the float depths */
void bubbleSort(Elem* a, int n)
{
  for(int i = n - 2; i >= 0; --i)
    for(int j = 0; j <= i; ++j)
      if(bitsf(a[j].d) >= bitsf(a[j + 1].d))
        std::swap(a[j], a[j + 1]);
}
/* insertionSort (oitCompositeDefines.glsl:94-109) */
void insertionSort(Elem* a, int L, Elem item)
{
  for(int i = 0; i < L; ++i)
    if(bitsf(item.d) < bitsf(a[i].d))
    {
      for(int j = L - 1; j > i; j--)
        a[j] = a[j - 1];
      a[i] = item;
      return;
    }
}
/* insertionSortTail (oitCompositeDefines.glsl:116-139) */
Elem insertionSortTail(Elem* a, int L, Elem item)
{
  Elem newlast = item;
  if(bitsf(item.d) < bitsf(a[L - 1].d))
    for(int i = 0; i < L; ++i)
      if(bitsf(item.d) < bitsf(a[i].d))
      {
        newlast = a[L - 1];
        for(int j = L - 1; j > i; j--)
          a[j] = a[j - 1];
        a[i] = item;
        break;
      }
  return newlast;
}
/* blend of the sorted array: per-sample coverage loop or plain (oitSimple.frag.glsl:138-167) */
void blendSorted(const OracleCtx* c, const Elem* a, int n, float colorSum[4])
{
  colorSum[0] = colorSum[1] = colorSum[2] = colorSum[3] = 0.f;
  if(c->coverage)
  {
    for(int s = 0; s < c->msaa; s++)
    {
      float sColor[4] = {0, 0, 0, 0};
      for(int i = 0; i < n; i++)
        if(a[i].m & (1u << s))
          doBlendPacked(sColor, a[i].c);
      for(int k = 0; k < 4; k++)
        colorSum[k] += sColor[k];
    }
    const float inv = 1.0f / (float)c->msaa;
    for(int k = 0; k < 4; k++)
      colorSum[k] *= inv;
  }
  else
    for(int i = 0; i < n; i++)
      doBlendPacked(colorSum, a[i].c);
}

void compositeInvocation(OracleCtx* c, uint32_t x, uint32_t y, uint32_t sampleID, float out[4])
{
  const size_t viewSize = (size_t)c->W * c->H;
  const size_t pix      = (size_t)y * c->W + x;
  const size_t ai       = ((size_t)sampleID * c->H + y) * c->W + x;
  const int    L        = (int)c->L;
  Elem         arr[32];
  switch(c->cfg.algorithm)
  {
    case OIT_SIMPLE:
    case OIT_SPINLOCK:
    case OIT_INTERLOCK: {
      /* oitSimple.frag.glsl:115-169 == oitSpinlock:153-207 == oitInterlock:176-230 */
      const uint32_t stride  = c->coverage ? 4 : 2;
      const size_t   listPos = viewSize * L * sampleID + pix;
      const int      n       = (int)std::min<uint32_t>((uint32_t)L, c->aux[ai]);
      for(int i = 0; i < n; i++)
      {
        const uint32_t* e = &c->abuf[(listPos + (size_t)i * viewSize) * stride];
        arr[i]            = Elem{e[0], e[1], c->coverage ? e[2] : 0u};
      }
      bubbleSort(arr, n);
      blendSorted(c, arr, n, out);
      break;
    }
    case OIT_LINKEDLIST: {
      /* oitLinkedList.frag.glsl:106-172 */
      int      n      = 0;
      uint32_t offset = c->aux[ai];
      while(offset != 0 && n < L)
      {
        const uint32_t* e = &c->abuf[(size_t)offset * 4];
        arr[n++]          = Elem{e[0], e[1], e[2]};
        offset            = e[3];
      }
      bubbleSort(arr, n);
      float tailColor[4] = {0, 0, 0, 0};
      while(offset != 0)
      {
        const uint32_t* e = &c->abuf[(size_t)offset * 4];
        const Elem      it{e[0], e[1], e[2]};
        if(c->cfg.tailBlend)
        {
          const Elem tail = insertionSortTail(arr, L, it);
          doBlendPacked(tailColor, tail.c);
        }
        else
          insertionSort(arr, L, it);
        offset = e[3];
      }
      blendSorted(c, arr, n, out);
      doBlend(out, tailColor);
      break;
    }
    case OIT_LOOP: {
      /* oitLoop.frag.glsl:191-222 */
      size_t listPos = viewSize * L * 2 * sampleID + pix;
      int    n       = 0;
      for(int i = 0; i < L; i++)
      {
        if(c->abuf[listPos + (size_t)i * viewSize] != 0xFFFFFFFFu)
          n++;
        else
          break;
      }
      listPos += viewSize * L;
      out[0] = out[1] = out[2] = out[3] = 0.f;
      for(int i = 0; i < n; i++)
        doBlendPacked(out, c->abuf[listPos + (size_t)i * viewSize]);
      break;
    }
    case OIT_LOOP64: {
      /* oitLoop64.frag.glsl:162-183 */
      const size_t listPos = viewSize * L * sampleID + pix;
      out[0] = out[1] = out[2] = out[3] = 0.f;
      for(int i = 0; i < L; i++)
      {
        const uint32_t* e = &c->abuf[(listPos + (size_t)i * viewSize) * 2];
        if(e[1] != 0xFFFFFFFFu)
          doBlendPacked(out, e[0]);
        else
          break;
      }
      break;
    }
    default: out[0] = out[1] = out[2] = out[3] = 0.f;
  }
}

/* std::mt19937 (the 1998 reference algorithm) + the generate_canonical<float, 24> of MSVC's STL as
   uniform_real_distribution<float> uses it: ONE 32-bit draw, float(x) / 2^32 (may round up to 1.0f, as MSVC's does) */
struct Mt19937
{
  uint32_t mt[624];
  int      idx;
  explicit Mt19937(uint32_t seed)
  {
    mt[0] = seed;
    for(int i = 1; i < 624; i++)
      mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next()
  {
    if(idx >= 624)
    {
      for(int k = 0; k < 624; k++)
      {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7FFFFFFFu);
        mt[k]            = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= y >> 18;
    return y;
  }
  float canonicalMsvc() { return (float)next() / 4294967296.0f; }
};

}  // namespace

/* ================================================================================================================ */
extern "C" {

float oracle_rand_canonical(uint64_t* state)
{
  /* std::default_random_engine == minstd_rand0 in libstdc++; uniform_real_distribution<float> draws one value and
     divides by float(range) (generate_canonical); known answers in tests */
  *state        = (*state * 16807ull) % 2147483647ull;
  const float r = (float)(*state - 1) / 2147483646.0f;
  return r >= 1.0f ? nextafterf(1.0f, 0.0f) : r;
}

int oracle_generate_scene_ex(const OracleConfig* cfg, int stdlib, float* vertices, uint32_t* indices);
int oracle_generate_scene(const OracleConfig* cfg, float* vertices, uint32_t* indices)
{
  return oracle_generate_scene_ex(cfg, 0, vertices, indices);
}

int oracle_scene_sizes(const OracleConfig* cfg, uint32_t* nVerts, uint32_t* nIndices, uint32_t* indicesPerObject)
{
  if(cfg->subdiv < 2 || cfg->numObjects < 1)
    return -1;
  const uint32_t sectors = cfg->subdiv * 2, stacks = cfg->subdiv;
  const uint32_t v = (sectors + 1) * (stacks + 1), t = sectors * (2 * stacks - 2);
  *nVerts          = v * cfg->numObjects;
  *indicesPerObject = t * 3;
  *nIndices         = t * 3 * cfg->numObjects;
  return 0;
}

/* initScene (main.cpp:334-391) with nvutils::createSphereUv(1, 2*subdiv, subdiv) restated (un-vendored nvpro_core2):
   stacks from +z pole to -z pole, (sectors+1) vertices per stack, two triangles per quad except at the poles */
int oracle_generate_scene_ex(const OracleConfig* cfg, int stdlib, float* vertices, uint32_t* indices)
{
  const int   sectors = cfg->subdiv * 2, stacks = cfg->subdiv;
  const float pi = 3.14159265358979323846f;
  const float sectorStep = 2.0f * pi / (float)sectors, stackStep = pi / (float)stacks;
  std::vector<float>    sp;
  std::vector<uint32_t> st;
  for(int i = 0; i <= stacks; ++i)
  {
    const float stackAngle = pi / 2.0f - (float)i * stackStep;
    const float xy = 1.0f * cosf(stackAngle), z = 1.0f * sinf(stackAngle);
    for(int j = 0; j <= sectors; ++j)
    {
      const float sectorAngle = (float)j * sectorStep;
      sp.push_back(xy * cosf(sectorAngle));
      sp.push_back(xy * sinf(sectorAngle));
      sp.push_back(z);
    }
  }
  for(int i = 0; i < stacks; ++i)
  {
    uint32_t k1 = i * (sectors + 1), k2 = k1 + sectors + 1;
    for(int j = 0; j < sectors; ++j, ++k1, ++k2)
    {
      if(i != 0)
      {
        st.push_back(k1);
        st.push_back(k2);
        st.push_back(k1 + 1);
      }
      if(i != stacks - 1)
      {
        st.push_back(k1 + 1);
        st.push_back(k2);
        st.push_back(k2 + 1);
      }
    }
  }
  const uint32_t nv = (uint32_t)sp.size() / 3, ni = (uint32_t)st.size();
  uint64_t       rng = 3625;  // main.cpp:350
  Mt19937        mt(3625u);
  /* std::default_random_engine is implementation-defined: minstd_rand0 in libstdc++ (stdlib 0), mt19937 in MSVC's STL
     (stdlib 1 -- the build that produced the screenshot in the reference's doc/, see tests/test_reference_screenshot.py) */
  auto oracle_rand_canonical = [&](uint64_t* state) -> float { return stdlib == 1 ? mt.canonicalMsvc() : ::oracle_rand_canonical(state); };
  for(int o = 0; o < cfg->numObjects; o++)
  {
    /* g++ and MSVC both evaluate the constructor arguments at main.cpp:356,366 right to left (g++ probed, SURVEY 8c;
       MSVC confirmed by the screenshot) */
    const float cz = oracle_rand_canonical(&rng), cy = oracle_rand_canonical(&rng), cx = oracle_rand_canonical(&rng);
    const float center[3] = {(cx - 0.5f) * 8.0f, (cy - 0.5f) * 8.0f, (cz - 0.5f) * 8.0f};
    float       radius    = 8.0f * 0.9f / 16;
    radius *= oracle_rand_canonical(&rng) * cfg->scaleWidth + cfg->scaleMin;
    const float ca = oracle_rand_canonical(&rng), cb = oracle_rand_canonical(&rng), cg = oracle_rand_canonical(&rng),
                cr = oracle_rand_canonical(&rng);
    const float color[4] = {cr * cr, cg * cg, cb * cb, ca};
    for(uint32_t v = 0; v < nv; v++)
    {
      float* d = &vertices[((size_t)o * nv + v) * 10];
      for(int k = 0; k < 3; k++)
      {
        d[k]     = sp[v * 3 + k] * radius + center[k];
        d[3 + k] = sp[v * 3 + k];
      }
      memcpy(d + 6, color, 16);
    }
    for(uint32_t i = 0; i < ni; i++)
      indices[(size_t)o * ni + i] = o * nv + st[i];
  }
  return 0;
}

/* CameraManipulator look-at + perspective (main.cpp:79-82,121-123,625-637): glm::lookAtRH, glm::perspectiveRH_ZO with
   the y axis flipped for Vulkan; the clip planes are harness parameters (nvpro_core2 defaults are not in the reference) */
void oracle_camera(uint32_t width, uint32_t height, float fovDeg, const float eye[3], const float center[3],
                   const float up[3], float zNear, float zFar, OracleSceneData* out)
{
  memset(out, 0, sizeof(*out));
  float f[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
  float fl   = 1.0f / sqrtf(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
  for(float& v : f)
    v *= fl;
  float s[3] = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
  float sl   = 1.0f / sqrtf(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
  for(float& v : s)
    v *= sl;
  const float u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
  float       V[16] = {s[0], u[0], -f[0], 0, s[1], u[1], -f[1], 0, s[2], u[2], -f[2], 0, 0, 0, 0, 1};
  V[12]            = -(s[0] * eye[0] + s[1] * eye[1] + s[2] * eye[2]);
  V[13]            = -(u[0] * eye[0] + u[1] * eye[1] + u[2] * eye[2]);
  V[14]            = (f[0] * eye[0] + f[1] * eye[1] + f[2] * eye[2]);
  const float aspect = (float)width / (float)height;
  const float th     = tanf(fovDeg * 0.01745329251994329577f / 2.0f);
  float       P[16]  = {0};
  P[0]               = 1.0f / (aspect * th);
  P[5]               = -(1.0f / th);
  P[10]              = zFar / (zNear - zFar);
  P[11]              = -1.0f;
  P[14]              = -(zFar * zNear) / (zFar - zNear);
  for(int c = 0; c < 4; c++)
    for(int r = 0; r < 4; r++)
    {
      float acc = 0.f;
      for(int k = 0; k < 4; k++)
        acc += P[k * 4 + r] * V[c * 4 + k];
      out->projViewMatrix[c * 4 + r] = acc;
    }
  memcpy(out->viewMatrix, V, sizeof(V));
  /* inverse transpose of a rigid view matrix: rotation part unchanged, translation moves to the last row */
  for(int c = 0; c < 3; c++)
    for(int r = 0; r < 3; r++)
      out->viewMatrixInverseTranspose[c * 4 + r] = V[c * 4 + r];
  for(int c = 0; c < 3; c++)
    out->viewMatrixInverseTranspose[c * 4 + 3] = -(V[c * 4 + 0] * V[12] + V[c * 4 + 1] * V[13] + V[c * 4 + 2] * V[14]);
  out->viewMatrixInverseTranspose[15] = 1.0f;
  out->viewport[0]                    = (int32_t)width;
  out->viewport[1]                    = (int32_t)height;
  out->viewport[2]                    = (int32_t)(width * height);
  out->alphaMin                       = 0.2f;  // main.cpp:126-127
  out->alphaWidth                     = 0.3f;
}

OracleCtx* oracle_create(const OracleConfig* cfg)
{
  if(cfg->algorithm >= NUM_ALGORITHMS || cfg->aaType >= NUM_AATYPES || cfg->oitLayers < 1 || cfg->oitLayers > 32
     || cfg->width == 0 || cfg->height == 0)
    return nullptr;
  OracleCtx* c = new OracleCtx();
  c->cfg       = *cfg;
  /* State::recomputeAntialiasingSettings (oit.h:88-115) */
  switch(cfg->aaType)
  {
    case AA_NONE: break;
    case AA_MSAA_4X: c->msaa = 4; break;
    case AA_SSAA_4X: c->msaa = 4; c->sampleShading = true; break;
    case AA_SUPER_4X: c->supersample = 2; break;
    case AA_MSAA_8X: c->msaa = 8; break;
    case AA_SSAA_8X: c->msaa = 8; c->sampleShading = true; break;
  }
  c->coverage = c->msaa > 1 && !c->sampleShading;
  c->W        = cfg->width * c->supersample;
  c->H        = cfg->height * c->supersample;
  c->L        = cfg->oitLayers;
  c->layers   = c->sampleShading ? c->msaa : 1;
  c->spos     = c->msaa == 1 ? SPOS1 : (c->msaa == 4 ? SPOS4 : SPOS8);
  /* createFrameImages (oit.cpp:84-163) */
  const size_t P = (size_t)c->W * c->H;
  size_t       words = 0;
  switch(cfg->algorithm)
  {
    case OIT_SIMPLE:
    case OIT_SPINLOCK:
    case OIT_INTERLOCK: words = P * c->L * (c->coverage ? 4 : 2); break;
    case OIT_LINKEDLIST:
      words       = P * (size_t)cfg->linkedListAllocatedPerElement * 4;
      c->capacity = (uint32_t)cfg->linkedListAllocatedPerElement * c->W * c->H;
      break;
    case OIT_LOOP: words = P * c->L * 2; break;
    case OIT_LOOP64: words = P * c->L * 2; break;
    default: break;
  }
  words *= c->layers;
  if(cfg->algorithm == OIT_LINKEDLIST)
    c->capacity *= c->layers;
  c->abuf.assign(words, 0u);
  if(cfg->algorithm != OIT_WEIGHTED)
    c->aux.assign(P * c->layers, 0u);
  if(cfg->algorithm == OIT_SPINLOCK)
    c->spin.assign(P * c->layers, 0u);
  if(cfg->algorithm == OIT_SPINLOCK || cfg->algorithm == OIT_INTERLOCK)
    c->adepth.assign(P * c->layers, 0u);
  if(cfg->algorithm == OIT_WEIGHTED)
  {
    c->wacc.assign(P * c->msaa * 4, 0);
    c->wrev.assign(P * c->msaa, 0);
  }
  c->color.assign(P * c->msaa, 0u);
  c->depth.assign(P * c->msaa, 1.0f);
  c->fin.assign((size_t)cfg->width * cfg->height, 0u);
  memset(&c->stats, 0, sizeof(c->stats));
  memset(&c->ubo, 0, sizeof(c->ubo));
  return c;
}
void oracle_destroy(OracleCtx* c) { delete c; }
int  oracle_set_threads(OracleCtx* c, int n)
{
  c->threads = n < 1 ? 1 : n;
  return 0;
}
int oracle_set_scene(OracleCtx* c, const float* vertices, uint32_t nVerts, const uint32_t* indices, uint32_t nIndices,
                     uint32_t indicesPerObject)
{
  if(indicesPerObject == 0 || indicesPerObject % 3 || nIndices % indicesPerObject)
    return -1;
  for(uint32_t i = 0; i < nIndices; i++)
    if(indices[i] >= nVerts)
      return -2;
  c->verts.assign(vertices, vertices + (size_t)nVerts * 10);
  c->idx.assign(indices, indices + nIndices);
  c->idxPerObj = indicesPerObject;
  return 0;
}
int oracle_set_scene_data(OracleCtx* c, const OracleSceneData* ubo)
{
  c->ubo             = *ubo;
  c->ubo.viewport[0] = (int32_t)c->W;  // updateUniformBuffer (main.cpp:628-637)
  c->ubo.viewport[1] = (int32_t)c->H;
  c->ubo.viewport[2] = (int32_t)(c->W * c->H);
  c->ubo.linkedListAllocatedPerElement = c->cfg.algorithm == OIT_LINKEDLIST ? c->capacity : c->L * c->layers;
  return 0;
}

/* clearTransparent* (oitRender.cpp:156-174,201-216,242-265,303-311,337-356) + render-pass clears (:89-91) */
int oracle_begin_frame(OracleCtx* c)
{
  memset(&c->stats, 0, sizeof(c->stats));
  const size_t viewSize = (size_t)c->W * c->H;
  switch(c->cfg.algorithm)
  {
    case OIT_SIMPLE: std::fill(c->aux.begin(), c->aux.end(), 0u); break;
    case OIT_LINKEDLIST:
      std::fill(c->aux.begin(), c->aux.end(), 0u);
      c->counter.store(0);
      break;
    case OIT_LOOP:
      for(uint32_t i = 0; i < c->layers; i++)
        std::fill_n(c->abuf.begin() + i * viewSize * c->L * 2, viewSize * c->L, 0xFFFFFFFFu);
      break;
    case OIT_LOOP64: std::fill(c->abuf.begin(), c->abuf.end(), 0xFFFFFFFFu); break;
    case OIT_SPINLOCK:
    case OIT_INTERLOCK:
      std::fill(c->adepth.begin(), c->adepth.end(), 0xFFFFFFFFu);
      std::fill(c->aux.begin(), c->aux.end(), 0u);
      std::fill(c->spin.begin(), c->spin.end(), 0u);
      break;
    default: break;
  }
  const float    clearLinear[4] = {0.2f, 0.2f, 0.2f, 0.2f};
  const uint32_t clearEnc       = encodeDst(clearLinear);
  std::fill(c->color.begin(), c->color.end(), clearEnc);
  std::fill(c->depth.begin(), c->depth.end(), 1.0f);
  if(c->cfg.algorithm == OIT_WEIGHTED)
  {
    /* WBOIT pass clears accum to 0 and reveal to 1 (oitRender.cpp:394-397) */
    std::fill(c->wacc.begin(), c->wacc.end(), (uint16_t)0);
    std::fill(c->wrev.begin(), c->wrev.end(), f2h(1.0f));
  }
  transformVertices(c);
  return 0;
}

static void splitObjects(const OracleCtx* c, uint32_t& numTransparent, uint32_t& numOpaque)
{
  /* oitRender.cpp:68-78 */
  const int numObjects = (int)(c->idx.size() / c->idxPerObj);
  int       nt         = (numObjects * c->cfg.percentTransparent) / 100;
  if(nt > numObjects)
    nt = numObjects;
  if(nt < 0)
    nt = 0;
  numTransparent = (uint32_t)nt;
  numOpaque      = (uint32_t)(numObjects - nt);
}

int oracle_draw_opaque(OracleCtx* c)
{
  uint32_t nt, no;
  splitObjects(c, nt, no);
  if(no > 0)
    drawRange(c, nt * c->idxPerObj, no * c->idxPerObj, 0);
  return 0;
}
int oracle_draw_transparent(OracleCtx* c)
{
  uint32_t nt, no;
  splitObjects(c, nt, no);
  c->stats.trianglesDrawn = (uint64_t)nt * c->idxPerObj / 3;
  if(nt == 0)
    return 0;
  if(c->cfg.algorithm == OIT_LOOP)
    drawRange(c, 0, nt * c->idxPerObj, 1);  // depth pass (oitRender.cpp:269-279)
  drawRange(c, 0, nt * c->idxPerObj, 2);
  c->counterOut      = c->counter.load();
  c->stats.llCounter = c->counterOut;
  return 0;
}

int oracle_composite(OracleCtx* c)
{
  const int S = c->msaa;
  if(c->cfg.algorithm == OIT_WEIGHTED)
  {
    /* oitWeighted.frag.glsl:98-109, per sample when msaa != 1 */
#pragma omp parallel for num_threads(c->threads) if(c->threads > 1)
    for(long p = 0; p < (long)((size_t)c->W * c->H); p++)
      for(int s = 0; s < S; s++)
      {
        const uint16_t* acc = &c->wacc[((size_t)p * S + s) * 4];
        const float     a3  = h2f(acc[3]);
        const float     den = a3 > 1e-5f ? a3 : 1e-5f;
        const float     src[4] = {h2f(acc[0]) / den, h2f(acc[1]) / den, h2f(acc[2]) / den, h2f(c->wrev[(size_t)p * S + s])};
        uint32_t&       d      = c->color[(size_t)p * S + s];
        d                      = ropWeightedComposite(d, src);
      }
    return 0;
  }
#pragma omp parallel for num_threads(c->threads) if(c->threads > 1)
  for(long y = 0; y < (long)c->H; y++)
    for(uint32_t x = 0; x < c->W; x++)
    {
      float out[4];
      if(c->sampleShading)
      {
        for(int s = 0; s < S; s++)
        {
          compositeInvocation(c, x, (uint32_t)y, (uint32_t)s, out);
          ropToSamples(c, x, (uint32_t)y, 1u << s, out);
        }
      }
      else
      {
        compositeInvocation(c, x, (uint32_t)y, 0, out);
        ropToSamples(c, x, (uint32_t)y, (1u << S) - 1u, out);
      }
    }
  return 0;
}

/* copyOffscreenToBackBuffer (main.cpp:645-774): box resolve of the S samples, or 2x2 LINEAR downsample, both done in
   linear space on the sRGB target (implementation-dependent in Vulkan; stated choice), then a raw copy */
int oracle_resolve(OracleCtx* c)
{
  const uint32_t w = c->cfg.width, h = c->cfg.height;
  const int      S = c->msaa, ss = c->supersample;
#pragma omp parallel for num_threads(c->threads) if(c->threads > 1)
  for(long y = 0; y < (long)h; y++)
    for(uint32_t x = 0; x < w; x++)
    {
      if(S == 1 && ss == 1)
      {
        c->fin[(size_t)y * w + x] = c->color[(size_t)y * w + x];
        continue;
      }
      float sum[4] = {0, 0, 0, 0};
      int   n      = 0;
      for(int dy = 0; dy < ss; dy++)
        for(int dx = 0; dx < ss; dx++)
          for(int s = 0; s < S; s++)
          {
            float d[4];
            decodeDst(c->color[(((size_t)y * ss + dy) * c->W + (x * ss + dx)) * S + s], d);
            for(int k = 0; k < 4; k++)
              sum[k] += d[k];
            n++;
          }
      const float inv = 1.0f / (float)n;
      for(int k = 0; k < 4; k++)
        sum[k] *= inv;
      c->fin[(size_t)y * w + x] = encodeDst(sum);
    }
  return 0;
}

int oracle_render(OracleCtx* c, const OracleSceneData* ubo)
{
  oracle_set_scene_data(c, ubo);
  oracle_begin_frame(c);
  oracle_draw_opaque(c);
  oracle_draw_transparent(c);
  oracle_composite(c);
  oracle_resolve(c);
  return 0;
}

int oracle_debug_invoke(OracleCtx* c, int pass, uint32_t x, uint32_t y, uint32_t sampleID, uint32_t coverageMask,
                        const float rgba[4], float depth, float viewZ, float outColor[4])
{
  if(x >= c->W || y >= c->H || sampleID >= c->layers)
    return -1;
  Frag f;
  f.x        = x;
  f.y        = y;
  f.sampleID = sampleID;
  f.mask     = coverageMask;
  memcpy(f.rgba, rgba, 16);
  f.z     = depth;
  f.viewz = viewZ;
  ThreadStats st;
  invoke(c, pass, f, st, outColor);
  c->stats.fragments += st.fragments;
  c->stats.fragmentsStored += st.stored;
  c->stats.fragmentsTail += st.tail;
  c->counterOut      = c->counter.load();
  c->stats.llCounter = c->counterOut;
  return 0;
}

uint32_t* oracle_abuffer(OracleCtx* c, size_t* n)
{
  *n = c->abuf.size();
  return c->abuf.data();
}
uint32_t* oracle_aux(OracleCtx* c, int which, size_t* n)
{
  switch(which)
  {
    case 0: *n = c->aux.size(); return c->aux.data();
    case 1: *n = c->spin.size(); return c->spin.data();
    case 2: *n = c->adepth.size(); return c->adepth.data();
    case 3: c->counterOut = c->counter.load(); *n = 1; return &c->counterOut;
  }
  *n = 0;
  return nullptr;
}
uint32_t* oracle_color_samples(OracleCtx* c, size_t* n)
{
  *n = c->color.size();
  return c->color.data();
}
float* oracle_depth_samples(OracleCtx* c, size_t* n)
{
  *n = c->depth.size();
  return c->depth.data();
}
uint16_t* oracle_weighted(OracleCtx* c, int which, size_t* n)
{
  if(which == 0)
  {
    *n = c->wacc.size();
    return c->wacc.data();
  }
  *n = c->wrev.size();
  return c->wrev.data();
}
uint32_t* oracle_final(OracleCtx* c, size_t* n)
{
  *n = c->fin.size();
  return c->fin.data();
}
int oracle_get_stats(OracleCtx* c, OracleStats* out)
{
  *out = c->stats;
  return 0;
}
int oracle_buffer_dims(OracleCtx* c, uint32_t* bufW, uint32_t* bufH, uint32_t* msaa, uint32_t* sampleShading)
{
  *bufW          = c->W;
  *bufH          = c->H;
  *msaa          = (uint32_t)c->msaa;
  *sampleShading = c->sampleShading ? 1u : 0u;
  return 0;
}

uint32_t oracle_srgb_encode8(float v) { return enc8(v); }
float    oracle_srgb_decode8(uint32_t v) { return T().dec[v & 255]; }
uint32_t oracle_pack_color(const float c[4]) { return packColor(c); }
uint16_t oracle_float_to_half(float f) { return f2h(f); }
float    oracle_half_to_float(uint16_t h) { return h2f(h); }
uint32_t oracle_rop_blend(uint32_t dst, const float src[4]) { return ropPremult(dst, src); }

}  // extern "C"
