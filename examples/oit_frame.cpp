// oit_frame.cpp -- renders one frame through the C++ mirror of the reference's Sample interface and prints the stats.
//   g++ -std=c++17 -Iinclude examples/oit_frame.cpp -Lvk_order_independent_transparency_b200 -loit_b200 -o build/oit_frame
//   LD_LIBRARY_PATH=vk_order_independent_transparency_b200 build/oit_frame [algorithm] [aaType] [width] [height]
#include <cstdio>
#include <cstdlib>

#include "oit_sample.hpp"

int main(int argc, char** argv)
{
  oitb200::State st;
  st.algorithm         = argc > 1 ? (uint32_t)atoi(argv[1]) : OIT_LINKEDLIST;
  st.aaType            = argc > 2 ? (uint32_t)atoi(argv[2]) : OIT_AA_NONE;
  const uint32_t width = argc > 3 ? (uint32_t)atoi(argv[3]) : 1280, height = argc > 4 ? (uint32_t)atoi(argv[4]) : 720;
  try
  {
    oitb200::Sample sample(st, width, height);
    sample.initScene();
    OitSceneData ubo;
    const float  eye[3] = {0.f, 0.f, 12.f}, center[3] = {0.f, 0.f, 0.f}, up[3] = {0.f, 1.f, 0.f};
    oit_default_camera(width, height, 45.f, eye, center, up, 0.1f, 100.f, &ubo);
    for(int i = 0; i < 3; i++)
      sample.onRender(ubo);
    const OitStats s   = sample.stats();
    const auto     img = sample.readColor();
    unsigned long long sum = 0;
    for(uint32_t v : img)
      sum += v;
    printf("fragments %llu stored %llu tail %llu  frame %.3f ms (geometry %.3f clear %.3f color %.3f composite %.3f resolve %.3f)  checksum %llu\n",
           (unsigned long long)s.fragments, (unsigned long long)s.fragmentsStored, (unsigned long long)s.fragmentsTail, s.msFrame,
           s.msGeometry, s.msClear, s.msColor, s.msComposite, s.msResolve, sum);
  }
  catch(const std::exception& e)
  {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
