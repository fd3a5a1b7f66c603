// oit_scene.cpp -- host-side harness inputs of the sample: the sphere-cloud scene (initScene, main.cpp:334-391) and the
// camera matrices (onAttach / updateUniformBuffer, main.cpp:79-82,121-123,625-637).  Pure host C++, no CUDA.
//
// The sphere mesh is nvutils::createSphereUv(1, 2*subdiv, subdiv) of the un-vendored nvpro_core2, restated: latitude
// rings from the +z pole to the -z pole with a duplicated seam vertex, two triangles per quad except at the poles.
// Random numbers are std::default_random_engine(3625) + uniform_real_distribution<float> as libstdc++ implements them
// (minstd_rand0, one draw per float), with g++'s right-to-left evaluation of the constructor arguments.
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/oit_b200.h"
#include "oit_scene.h"

namespace {

struct MinStd
{
  uint64_t s;
  float    next()
  {
    s             = (s * 16807ull) % 2147483647ull;
    const float r = (float)(s - 1) / 2147483646.0f;
    return r < 1.0f ? r : std::nextafterf(1.0f, 0.0f);
  }
};

// std::mt19937 + MSVC's generate_canonical<float, 24> (one 32-bit draw / 2^32, may round up to 1.0f): what
// std::default_random_engine and uniform_real_distribution<float> are in MSVC's STL (OIT_STDLIB_MSVC)
struct Mt19937
{
  uint32_t state[624];
  int      pos = 624;
  explicit Mt19937(uint32_t seed)
  {
    state[0] = seed;
    for(uint32_t i = 1; i < 624; i++)
      state[i] = 1812433253u * (state[i - 1] ^ (state[i - 1] >> 30)) + i;
  }
  void twist()
  {
    for(int i = 0; i < 624; i++)
    {
      const uint32_t mixed = (state[i] & 0x80000000u) | (state[(i + 1) % 624] & 0x7FFFFFFFu);
      state[i]             = state[(i + 397) % 624] ^ (mixed >> 1) ^ ((mixed & 1u) ? 0x9908B0DFu : 0u);
    }
    pos = 0;
  }
  float next()
  {
    if(pos >= 624)
      twist();
    uint32_t v = state[pos++];
    v ^= v >> 11;
    v ^= (v << 7) & 0x9D2C5680u;
    v ^= (v << 15) & 0xEFC60000u;
    v ^= v >> 18;
    return (float)v / 4294967296.0f;
  }
};

struct UnitSphere
{
  std::vector<float>    pos;  // xyz per vertex (== normal)
  std::vector<uint32_t> tri;
};

UnitSphere makeUnitSphere(int subdiv)
{
  UnitSphere  m;
  const int   sectors = 2 * subdiv, stacks = subdiv;
  const float kPi        = 3.14159265358979323846f;
  const float dSector = 2.0f * kPi / (float)sectors, dStack = kPi / (float)stacks;
  m.pos.reserve((size_t)(sectors + 1) * (stacks + 1) * 3);
  for(int ring = 0; ring <= stacks; ring++)
  {
    const float lat   = kPi / 2.0f - (float)ring * dStack;
    const float rxy   = 1.0f * cosf(lat);
    const float z     = 1.0f * sinf(lat);
    for(int seg = 0; seg <= sectors; seg++)
    {
      const float lon = (float)seg * dSector;
      m.pos.push_back(rxy * cosf(lon));
      m.pos.push_back(rxy * sinf(lon));
      m.pos.push_back(z);
    }
  }
  for(int ring = 0; ring < stacks; ring++)
    for(int seg = 0; seg < sectors; seg++)
    {
      const uint32_t upper = (uint32_t)(ring * (sectors + 1) + seg), lower = upper + (uint32_t)sectors + 1u;
      if(ring != 0)
        m.tri.insert(m.tri.end(), {upper, lower, upper + 1u});
      if(ring != stacks - 1)
        m.tri.insert(m.tri.end(), {upper + 1u, lower, lower + 1u});
    }
  return m;
}

}  // namespace

extern "C" {

int oit_scene_sizes(const OitConfig* cfg, uint32_t* nVerts, uint32_t* nIndices, uint32_t* indicesPerObject)
{
  if(!cfg || cfg->subdiv < 2 || cfg->subdiv > 1024 || cfg->numObjects < 1)
    return OIT_ERR_INVALID_ARG;
  const uint64_t sectors = 2ull * cfg->subdiv, stacks = (uint64_t)cfg->subdiv;
  const uint64_t v = (sectors + 1) * (stacks + 1), i = sectors * (2 * stacks - 2) * 3;
  if(v * cfg->numObjects > 0xFFFFFFFFull || i * cfg->numObjects > 0xFFFFFFFFull)
    return OIT_ERR_INVALID_ARG;
  if(nVerts)
    *nVerts = (uint32_t)(v * cfg->numObjects);
  if(nIndices)
    *nIndices = (uint32_t)(i * cfg->numObjects);
  if(indicesPerObject)
    *indicesPerObject = (uint32_t)i;
  return OIT_OK;
}

int oit_generate_spheres(const OitConfig* cfg, OitSphere* spheres)
{
  if(!spheres || oit_scene_sizes(cfg, nullptr, nullptr, nullptr) != OIT_OK)
    return OIT_ERR_INVALID_ARG;
  MinStd      minstd{3625};  // main.cpp:350
  Mt19937     mt(3625u);
  const bool  msvc = OIT_CFG_SCENE_STDLIB(cfg) == OIT_STDLIB_MSVC;
  struct
  {
    MinStd&  a;
    Mt19937& b;
    bool     useB;
    float    next() { return useB ? b.next() : a.next(); }
  } rng{minstd, mt, msvc};
  const float kGlobalScale = 8.0f, kGrid = 16.0f;  // GLOBAL_SCALE, GRID_SIZE (main.cpp:54-55)
  for(int obj = 0; obj < cfg->numObjects; obj++)
  {
    OitSphere& o = spheres[obj];
    // glm::vec3 center(u(), u(), u()): z is drawn first, x last
    o.center[2] = rng.next();
    o.center[1] = rng.next();
    o.center[0] = rng.next();
    for(float& v : o.center)
      v = (v - 0.5f) * kGlobalScale;
    float radius = kGlobalScale * 0.9f / kGrid;
    radius *= rng.next() * cfg->scaleWidth + cfg->scaleMin;
    o.radius = radius;
    // glm::vec4 color(u(), u(), u(), u()): alpha first, red last; rgb squared (main.cpp:366-369)
    o.color[3] = rng.next();
    o.color[2] = rng.next();
    o.color[1] = rng.next();
    o.color[0] = rng.next();
    o.color[0] *= o.color[0];
    o.color[1] *= o.color[1];
    o.color[2] *= o.color[2];
  }
  return OIT_OK;
}

int oit_generate_scene(const OitConfig* cfg, void* vertices, uint32_t* indices)
{
  uint32_t nv, ni, ipo;
  if(!vertices || !indices || oit_scene_sizes(cfg, &nv, &ni, &ipo) != OIT_OK)
    return OIT_ERR_INVALID_ARG;
  const UnitSphere       sphere = makeUnitSphere(cfg->subdiv);
  const uint32_t         vPer   = (uint32_t)(sphere.pos.size() / 3);
  float*                 out    = static_cast<float*>(vertices);
  std::vector<OitSphere> table((size_t)cfg->numObjects);
  oit_generate_spheres(cfg, table.data());
  for(int obj = 0; obj < cfg->numObjects; obj++)
  {
    const OitSphere& o   = table[(size_t)obj];
    float*           dst = out + (size_t)obj * vPer * 10;
    for(uint32_t v = 0; v < vPer; v++, dst += 10)
    {
      const float* p = &sphere.pos[(size_t)v * 3];
      dst[0]         = p[0] * o.radius + o.center[0];
      dst[1]         = p[1] * o.radius + o.center[1];
      dst[2]         = p[2] * o.radius + o.center[2];
      dst[3]         = p[0];
      dst[4]         = p[1];
      dst[5]         = p[2];
      memcpy(dst + 6, o.color, sizeof(o.color));
    }
    uint32_t* idst = indices + (size_t)obj * ipo;
    for(uint32_t i = 0; i < ipo; i++)
      idst[i] = (uint32_t)obj * vPer + sphere.tri[i];
  }
  return OIT_OK;
}

// glm::lookAtRH + glm::perspectiveRH_ZO with the y axis flipped for Vulkan, proj * view in float like glm.
int oit_default_camera(uint32_t width, uint32_t height, float fovDeg, const float eye[3], const float center[3],
                       const float up[3], float zNear, float zFar, OitSceneData* out)
{
  if(!out || !eye || !center || !up || width == 0 || height == 0 || !(zFar > zNear) || !(zNear > 0.f))
    return OIT_ERR_INVALID_ARG;
  memset(out, 0, sizeof(*out));
  auto normalize3 = [](float v[3]) {
    const float inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] *= inv;
    v[1] *= inv;
    v[2] *= inv;
  };
  auto dot3 = [](const float a[3], const float b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
  float fwd[3] = {center[0] - eye[0], center[1] - eye[1], center[2] - eye[2]};
  normalize3(fwd);
  float side[3] = {fwd[1] * up[2] - fwd[2] * up[1], fwd[2] * up[0] - fwd[0] * up[2], fwd[0] * up[1] - fwd[1] * up[0]};
  normalize3(side);
  const float upv[3] = {side[1] * fwd[2] - side[2] * fwd[1], side[2] * fwd[0] - side[0] * fwd[2], side[0] * fwd[1] - side[1] * fwd[0]};
  float*      V      = out->viewMatrix;
  for(int col = 0; col < 3; col++)
  {
    V[col * 4 + 0] = side[col];
    V[col * 4 + 1] = upv[col];
    V[col * 4 + 2] = -fwd[col];
    V[col * 4 + 3] = 0.f;
  }
  V[12] = -dot3(side, eye);
  V[13] = -dot3(upv, eye);
  V[14] = dot3(fwd, eye);
  V[15] = 1.f;
  float       Pm[16] = {0};
  const float aspect = (float)width / (float)height;
  const float tanHalf = tanf(fovDeg * 0.01745329251994329577f / 2.0f);
  Pm[0]               = 1.0f / (aspect * tanHalf);
  Pm[5]               = -(1.0f / tanHalf);
  Pm[10]              = zFar / (zNear - zFar);
  Pm[11]              = -1.0f;
  Pm[14]              = -(zFar * zNear) / (zFar - zNear);
  for(int col = 0; col < 4; col++)
    for(int row = 0; row < 4; row++)
    {
      float acc = 0.f;
      for(int k = 0; k < 4; k++)
        acc += Pm[k * 4 + row] * V[col * 4 + k];
      out->projViewMatrix[col * 4 + row] = acc;
    }
  // inverse transpose of the rigid view matrix
  float* IT = out->viewMatrixInverseTranspose;
  for(int col = 0; col < 3; col++)
  {
    for(int row = 0; row < 3; row++)
      IT[col * 4 + row] = V[col * 4 + row];
    IT[col * 4 + 3] = -(V[col * 4 + 0] * V[12] + V[col * 4 + 1] * V[13] + V[col * 4 + 2] * V[14]);
  }
  IT[15]           = 1.0f;
  out->viewport[0] = (int32_t)width;
  out->viewport[1] = (int32_t)height;
  out->viewport[2] = (int32_t)(width * height);
  out->alphaMin    = 0.2f;  // main.cpp:126-127
  out->alphaWidth  = 0.3f;
  return OIT_OK;
}

}  // extern "C"

namespace oit {
void unitSphereTemplate(int subdiv, std::vector<float>& pos, std::vector<uint32_t>& tri)
{
  UnitSphere m = makeUnitSphere(subdiv);
  pos.swap(m.pos);
  tri.swap(m.tri);
}
}  // namespace oit
