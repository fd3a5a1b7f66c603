// oit_sample.hpp -- header-only C++ mirror of the reference's hot-path interface over the C ABI of liboit_b200.so.
//
// Same method names, argument meaning and error behaviour as `Sample` (oit.h:379-425, oitRender.cpp:28-425):
//   onRender, clearTransparent{Simple,LinkedList,Loop,Loop64,Lock}, drawTransparent{Simple,LinkedList,Loop,Loop64,Lock,
//   Weighted}, copyOffscreenToBackBuffer; `State` has the reference's fields and defaults (oit.h:64-116).
// The reference aborts through NVVK_CHECK / assert; this mirror throws std::runtime_error carrying oit_last_error().
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "oit_b200.h"

namespace oitb200 {

struct State  // oit.h:64-82
{
  uint32_t algorithm                     = OIT_SPINLOCK;
  uint32_t oitLayers                     = 8;
  int32_t  linkedListAllocatedPerElement = 10;
  int32_t  percentTransparent            = 100;
  bool     tailBlend                     = true;
  bool     interlockIsOrdered            = true;
  int32_t  numObjects                    = 1024;
  int32_t  subdiv                        = 16;
  float    scaleMin                      = 0.1f;
  float    scaleWidth                    = 0.9f;
  uint32_t aaType                        = OIT_AA_NONE;
  // implicitly set by aaType (oit.h:84-115)
  int  msaa          = 1;
  bool sampleShading = false;
  int  supersample   = 1;
  bool coverageShading() const { return msaa > 1 && !sampleShading; }
  void recomputeAntialiasingSettings()
  {
    sampleShading = false;
    supersample   = 1;
    switch(aaType)
    {
      case OIT_AA_NONE: msaa = 1; break;
      case OIT_AA_MSAA_4X: msaa = 4; break;
      case OIT_AA_SSAA_4X: msaa = 4; sampleShading = true; break;
      case OIT_AA_SUPER_4X: msaa = 1; supersample = 2; break;
      case OIT_AA_MSAA_8X: msaa = 8; break;
      case OIT_AA_SSAA_8X: msaa = 8; sampleShading = true; break;
      default: throw std::runtime_error("Antialiasing mode not implemented!");
    }
  }
};

class Sample
{
public:
  Sample(const State& state, uint32_t width, uint32_t height, int device = 0, uint32_t bandCount = 1, uint32_t bandIndex = 0,
         uint32_t stripRows = 32)
      : m_state(state)
  {
    m_state.recomputeAntialiasingSettings();
    OitConfig cfg;
    oit_default_config(&cfg);
    cfg.algorithm                     = m_state.algorithm;
    cfg.oitLayers                     = m_state.oitLayers;
    cfg.linkedListAllocatedPerElement = m_state.linkedListAllocatedPerElement;
    cfg.percentTransparent            = m_state.percentTransparent;
    cfg.tailBlend                     = m_state.tailBlend ? 1u : 0u;
    cfg.interlockIsOrdered            = m_state.interlockIsOrdered ? 1u : 0u;
    cfg.numObjects                    = m_state.numObjects;
    cfg.subdiv                        = m_state.subdiv;
    cfg.scaleMin                      = m_state.scaleMin;
    cfg.scaleWidth                    = m_state.scaleWidth;
    cfg.aaType                        = m_state.aaType;
    cfg.width                         = width;
    cfg.height                        = height;
    cfg.device                        = device;
    cfg.bandCount                     = bandCount;
    cfg.bandIndex                     = bandIndex;
    cfg.stripRows                     = stripRows;
    m_cfg                             = cfg;
    if(oit_create(&cfg, &m_ctx) != OIT_OK)
      throw std::runtime_error(std::string("oit_create: ") + oit_last_error(nullptr));
    oit_get_dims(m_ctx, nullptr, nullptr, nullptr, nullptr, &m_localRows);
  }
  ~Sample() { oit_destroy(m_ctx); }
  Sample(const Sample&)            = delete;
  Sample& operator=(const Sample&) = delete;

  // initScene (main.cpp:334-417): the sample's sphere cloud for m_state, generated on the host and uploaded
  void initScene()
  {
    uint32_t nv = 0, ni = 0, ipo = 0;
    check(oit_scene_sizes(&m_cfg, &nv, &ni, &ipo));
    std::vector<float>    verts((size_t)nv * 10);
    std::vector<uint32_t> idx(ni);
    check(oit_generate_scene(&m_cfg, verts.data(), idx.data()));
    check(oit_set_scene(m_ctx, verts.data(), nv, idx.data(), ni, ipo));
  }
  void setScene(const void* vertices, uint32_t nVerts, const uint32_t* indices, uint32_t nIndices, uint32_t indicesPerObject)
  {
    check(oit_set_scene(m_ctx, vertices, nVerts, indices, nIndices, indicesPerObject));
  }
  // the same scene from its 32-byte-per-object table; the mesh is flattened on the device
  void initSceneInstanced()
  {
    std::vector<OitSphere> table((size_t)m_cfg.numObjects);
    check(oit_generate_spheres(&m_cfg, table.data()));
    check(oit_set_scene_spheres(m_ctx, table.data(), (uint32_t)table.size(), m_cfg.subdiv));
  }

  void onRender(const OitSceneData& ubo) { check(oit_render(m_ctx, &ubo)); }  // oitRender.cpp:28-154
  void updateUniformBuffer(const OitSceneData& ubo) { check(oit_set_scene_data(m_ctx, &ubo)); }

  void clearTransparentSimple() { clear(OIT_SIMPLE, OIT_SIMPLE); }
  void clearTransparentLinkedList() { clear(OIT_LINKEDLIST, OIT_LINKEDLIST); }
  void clearTransparentLoop() { clear(OIT_LOOP, OIT_LOOP); }
  void clearTransparentLoop64() { clear(OIT_LOOP64, OIT_LOOP64); }
  void clearTransparentLock(bool useInterlock) { clear(useInterlock ? OIT_INTERLOCK : OIT_SPINLOCK, useInterlock ? OIT_INTERLOCK : OIT_SPINLOCK); }
  void drawOpaque() { check(oit_draw_opaque(m_ctx)); }
  void drawTransparentSimple() { draw(OIT_SIMPLE); }
  void drawTransparentLinkedList() { draw(OIT_LINKEDLIST); }
  void drawTransparentLoop() { draw(OIT_LOOP); }
  void drawTransparentLoop64() { draw(OIT_LOOP64); }
  void drawTransparentLock(bool useInterlock) { draw(useInterlock ? OIT_INTERLOCK : OIT_SPINLOCK); }
  void drawTransparentWeighted() { draw(OIT_WEIGHTED); }
  void copyOffscreenToBackBuffer()
  {
    check(oit_resolve(m_ctx));
    check(oit_synchronize(m_ctx));
  }

  // m_viewportImage rows owned by this band: BGRA8, sRGB-encoded bytes
  std::vector<uint32_t> readColor()
  {
    std::vector<uint32_t> out((size_t)m_cfg.width * m_localRows);
    if(!out.empty())
      check(oit_read_color(m_ctx, out.data(), out.size() * 4));
    return out;
  }
  OitStats stats()
  {
    OitStats s;
    check(oit_get_stats(m_ctx, &s));
    return s;
  }
  OitCtx*      handle() { return m_ctx; }
  const State& state() const { return m_state; }
  uint32_t     localRows() const { return m_localRows; }

private:
  void check(int r)
  {
    if(r != OIT_OK)
      throw std::runtime_error(std::string("liboit_b200: ") + oit_last_error(m_ctx));
  }
  void clear(uint32_t a, uint32_t b)
  {
    if(m_state.algorithm != a && m_state.algorithm != b)
      throw std::runtime_error("Algorithm case not called in switch statement!");  // oitRender.cpp:65
    check(oit_begin_frame(m_ctx));
  }
  void draw(uint32_t a)
  {
    if(m_state.algorithm != a)
      throw std::runtime_error("Algorithm case not called in switch statement!");  // oitRender.cpp:147
    check(oit_draw_transparent(m_ctx));
    check(oit_composite(m_ctx));
  }
  State     m_state;
  OitConfig m_cfg{};
  OitCtx*   m_ctx       = nullptr;
  uint32_t  m_localRows = 0;
};

}  // namespace oitb200
