// oit_scene.h -- the unit UV sphere of the sample's scene generator (oit_scene.cpp), shared with the instanced scene input
#pragma once
#include <cstdint>
#include <vector>

namespace oit {
// nvutils::createSphereUv(1, 2 * subdiv, subdiv) restated: xyz per vertex (== normal), triangle list
void unitSphereTemplate(int subdiv, std::vector<float>& pos, std::vector<uint32_t>& tri);
}  // namespace oit
