// oit_clip.cuh -- near-plane clipping (the fixed-function clipper in front of the rasteriser: 0 <= z_clip).
//
// A triangle with a vertex behind the near plane has no valid post-projection position for that vertex (TVert::x ==
// INT32_MIN).  The binning kernels cut it against z_clip >= 0 in clip space into one triangle (two vertices behind) or a
// quad = two triangles (one vertex behind) and store every piece that owns tiles in the frame's clip table (ClipEntry:
// the piece's post-projection vertices and vertex records); the piece's (tile, triangle) pairs carry the entry's index
// (PAIR_CLIPPED).  The raster kernel reads a piece exactly like a triangle, only through the entry instead of the vertex
// buffer, so the clipper itself never runs in the frame kernel.  Triangles that reach the near plane are rare (the
// default camera has none); when they occur they are few and large.
//
// Rules (DESIGN.md "Arithmetic contract"; the CPU checker of the test suite states the same ones):
//  * a vertex is inside iff z_clip >= 0; clip coordinates are recomputed with the vertex stage's own fma chain;
//  * a new vertex lies on an edge from an INSIDE vertex P to an OUTSIDE vertex Q (always in that direction, so two
//    triangles that share the edge produce the identical vertex): t = zP / (zP - zQ), x / y / w / view-z = fma(t, Q - P, P),
//    z_clip := 0; then the vertex stage's perspective divide, viewport transform and snapping;
//  * one vertex outside (k; a = k+1, b = k+2): A' on a->k, B' on b->k, pieces (A', a, b) and (A', b, B');
//    two outside (inside vertex a; b = a+1, c = a+2): P on a->b, Q on a->c, piece (a, P, Q) -- orientation kept;
//  * the vertex record (normal, colour) of a new vertex = fma(t, attr[Q] - attr[P], attr[P]): linear in clip space;
//  * if any resulting vertex is not representable (w <= 0, guard band, z outside [0, 1]) the whole triangle stays
//    rejected, as do triangles whose vertices are invalid for any other reason than the near plane (far plane, guard band).
#pragma once
#include "oit_device.cuh"

namespace oit {

// the vertex stage after the matrix product: perspective divide, viewport transform, guard band, snapping to 1/256 px
__device__ __forceinline__ TVert finishVertex(const float clip[4], float hw, float hh)
{
  TVert t;
  t.x     = INT32_MIN;
  t.y     = 0;
  t.z     = 0.f;
  t.invw  = 0.f;
  if(clip[3] > 0.f && clip[3] < __int_as_float(0x7f800000))
  {
    const float invw = __fdiv_rn(1.0f, clip[3]);
    const float nx = __fmul_rn(clip[0], invw), ny = __fmul_rn(clip[1], invw), nz = __fmul_rn(clip[2], invw);
    const float xs = __fmaf_rn(nx, hw, hw), ys = __fmaf_rn(ny, hh, hh);
    if(fabsf(xs) < GUARD_BAND_PX && fabsf(ys) < GUARD_BAND_PX && nz >= 0.f && nz <= 1.f)
    {
      t.x    = __float2int_rn(__fmul_rn(xs, 256.0f));
      t.y    = __float2int_rn(__fmul_rn(ys, 256.0f));
      t.z    = nz;
      t.invw = invw;
    }
  }
  return t;
}

// what the clipper reads, passed BY VALUE: handing the kernel's FrameParams to an out-of-line function by reference would
// force a local-memory copy of the whole parameter block and slow every access to it down
struct ClipInput
{
  const float*     verts;
  const DeviceUbo* ubo;
  int              W, H;
};
__device__ __forceinline__ ClipInput clipInput(const FrameParams& p) { return ClipInput{p.verts, p.ubo, p.W, p.H}; }

__device__ __forceinline__ void clipSpaceVertex(const ClipInput& p, uint32_t index, float clip[4], float& viewz)
{
  const float* M  = p.ubo->projView;
  const float* V  = p.ubo->view;
  const float* v  = p.verts + (size_t)index * 10;
  const float  px = v[0], py = v[1], pz = v[2];
#pragma unroll
  for(int r = 0; r < 4; r++)
    clip[r] = __fmaf_rn(M[0 + r], px, __fmaf_rn(M[4 + r], py, __fmaf_rn(M[8 + r], pz, M[12 + r])));
  viewz = __fmaf_rn(V[2], px, __fmaf_rn(V[6], py, __fmaf_rn(V[10], pz, V[14])));
}

struct ClipVert
{
  TVert v;
  float viewz;
  int   i, j;  // attribute source: original vertex i (t == 0), or the point t of the way from i to j
  float t;
};
struct ClipResult
{
  int      count;  // sub-triangles: 0 (stays rejected), 1 or 2
  ClipVert v[2][3];
};

// ix: the triangle's three vertex indices in index-buffer order
static __device__ __noinline__ ClipResult clipTriangleNear(const ClipInput p, uint32_t ix0, uint32_t ix1, uint32_t ix2)
{
  ClipResult     r;
  r.count = 0;
  const uint32_t ix[3] = {ix0, ix1, ix2};
  const float    hw = 0.5f * (float)p.W, hh = 0.5f * (float)p.H;
  float          clip[3][4], vz[3];
  int            nIn = 0, firstOut = -1, firstIn = -1;
#pragma unroll 1
  for(int k = 0; k < 3; k++)
  {
    clipSpaceVertex(p, ix[k], clip[k], vz[k]);
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
    return r;
  auto original = [&](int k) {
    ClipVert c;
    c.v     = finishVertex(clip[k], hw, hh);
    c.viewz = vz[k];
    c.i = k;
    c.j = k;
    c.t = 0.f;
    return c;
  };
  auto cut = [&](int in, int out) {  // the point of the edge in -> out on the plane z_clip = 0
    const float zP = clip[in][2], zQ = clip[out][2];
    const float t  = __fdiv_rn(zP, __fsub_rn(zP, zQ));
    float       c4[4];
    c4[0] = __fmaf_rn(t, __fsub_rn(clip[out][0], clip[in][0]), clip[in][0]);
    c4[1] = __fmaf_rn(t, __fsub_rn(clip[out][1], clip[in][1]), clip[in][1]);
    c4[2] = 0.f;
    c4[3] = __fmaf_rn(t, __fsub_rn(clip[out][3], clip[in][3]), clip[in][3]);
    ClipVert c;
    c.v     = finishVertex(c4, hw, hh);
    c.viewz = __fmaf_rn(t, __fsub_rn(vz[out], vz[in]), vz[in]);
    c.i = in;
    c.j = out;
    c.t = t;
    return c;
  };
  if(nIn == 2)
  {
    const int k = firstOut, a = (k + 1) % 3, b = (k + 2) % 3;
    const ClipVert A = cut(a, k), B = cut(b, k), va = original(a), vb = original(b);
    r.v[0][0] = A;
    r.v[0][1] = va;
    r.v[0][2] = vb;
    r.v[1][0] = A;
    r.v[1][1] = vb;
    r.v[1][2] = B;
    r.count   = 2;
  }
  else
  {
    const int a = firstIn, b = (a + 1) % 3, c = (a + 2) % 3;
    r.v[0][0] = original(a);
    r.v[0][1] = cut(a, b);
    r.v[0][2] = cut(a, c);
    r.count   = 1;
  }
  for(int s = 0; s < r.count; s++)
    for(int k = 0; k < 3; k++)
      if(r.v[s][k].v.x == INT32_MIN)
        r.count = 0;  // something is not representable: the triangle stays rejected
  return r;
}

}  // namespace oit
