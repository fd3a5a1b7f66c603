// oit_gather.cu -- the band gather of the split-frame mode, inside the library: ONE ncclAllGather of the resolved strips
// (in place: every rank resolves straight into its slice of the gather buffer) followed by the row interleave, both
// enqueued on the context's stream so that they become the last nodes of the captured frame graph.
//
// NCCL is loaded with dlopen (libnccl.so.2 -- the copy torch already mapped if the host process imported torch, else
// the system one), so liboit_b200.so has no link-time dependency on it and single-GPU hosts never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>

#include "oit_internal.h"

namespace oit {

namespace {
struct NcclApi
{
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*)                                                               = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int)                                        = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)          = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t)                                                                  = nullptr;
  const char* (*GetErrorString)(ncclResult_t)                                                              = nullptr;
};

NcclApi* ncclApi(std::string& err)
{
  static NcclApi api;
  static bool    tried = false;
  if(!tried)
  {
    tried = true;
    for(const char* name : {"libnccl.so.2", "libnccl.so"})
      if((api.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL)) != nullptr)
        break;
    if(api.handle)
    {
      api.GetUniqueId    = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(api.handle, "ncclGetUniqueId"));
      api.CommInitRank   = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(api.handle, "ncclCommInitRank"));
      api.AllGather      = reinterpret_cast<decltype(api.AllGather)>(dlsym(api.handle, "ncclAllGather"));
      api.CommDestroy    = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(api.handle, "ncclCommDestroy"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(api.handle, "ncclGetErrorString"));
    }
  }
  if(!api.handle || !api.GetUniqueId || !api.CommInitRank || !api.AllGather || !api.CommDestroy)
  {
    err = "NCCL (libnccl.so.2) could not be loaded";
    return nullptr;
  }
  return &api;
}
}  // namespace

struct BandGatherState
{
  ncclComm_t comm = nullptr;
  int        rank = 0, world = 1;
};

int gatherUniqueId(void* id128, std::string& err)
{
  NcclApi* api = ncclApi(err);
  if(!api)
    return OIT_ERR_UNSUPPORTED;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclUniqueId id;
  const ncclResult_t r = api->GetUniqueId(&id);
  if(r != ncclSuccess)
  {
    err = std::string("ncclGetUniqueId: ") + (api->GetErrorString ? api->GetErrorString(r) : "error");
    return OIT_ERR_CUDA;
  }
  memcpy(id128, &id, 128);
  return OIT_OK;
}

BandGatherState* gatherCreate(const void* id128, int rank, int world, std::string& err)
{
  NcclApi* api = ncclApi(err);
  if(!api)
    return nullptr;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  BandGatherState*   g = new BandGatherState();
  g->rank              = rank;
  g->world             = world;
  const ncclResult_t r = api->CommInitRank(&g->comm, world, id, rank);
  if(r != ncclSuccess)
  {
    err = std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "error");
    delete g;
    return nullptr;
  }
  return g;
}

void gatherDestroy(BandGatherState* g)
{
  if(!g)
    return;
  std::string err;
  NcclApi*    api = ncclApi(err);
  if(api && g->comm)
    api->CommDestroy(g->comm);
  delete g;
}

// frame[y][x] = gathered[band(y)][localRow(y)][x]; band k owns the strips k, k + G, ... of stripRows output rows.
// Every band's slice is padRows rows + ONE metadata row whose first word pair is the band's overflow flag of this frame
// (a pair / clip buffer was too small): their OR goes to stats[STAT_OVERFLOW_ANY], so that every band takes the same
// decision to render the frame again.
__global__ void __launch_bounds__(256) k_interleave_rows(const uint4* __restrict__ gathered, uint4* __restrict__ frame, int quadsPerRow, int H,
                                                         int stripRows, int G, int padRows, unsigned long long* __restrict__ stats)
{
  if(blockIdx.x == 0 && threadIdx.x == 0)
  {
    unsigned long long any = 0;
    for(int b = 0; b < G; b++)
      any |= *reinterpret_cast<const unsigned long long*>(gathered + ((size_t)b * (padRows + 1) + padRows) * quadsPerRow);
    stats[STAT_OVERFLOW_ANY] = any ? 1ull : 0ull;
  }
  const size_t total = (size_t)quadsPerRow * H;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    const int y = (int)(i / quadsPerRow), q = (int)(i - (size_t)y * quadsPerRow);
    const int strip = y / stripRows, band = strip % G;
    const int local = (strip / G) * stripRows + (y - strip * stripRows);
    frame[i]        = gathered[((size_t)band * (padRows + 1) + local) * quadsPerRow + q];
  }
}

// all-gather (in place) + interleave; returns the number of own kernels launched, < 0 on error
int gatherLaunch(BandGatherState* g, uint32_t* gathered, uint32_t* frame, int W, int H, int stripRows, int padRows, unsigned long long* stats,
                 cudaStream_t s, std::string& err)
{
  NcclApi* api = ncclApi(err);
  if(!api)
    return OIT_ERR_UNSUPPORTED;
  const size_t count = (size_t)(padRows + 1) * W;  // + the metadata row
  if(cudaMemcpyAsync(gathered + (size_t)g->rank * count + (size_t)padRows * W, stats + STAT_OVERFLOW, sizeof(unsigned long long),
                     cudaMemcpyDeviceToDevice, s) != cudaSuccess)
  {
    err = "band gather: metadata copy failed";
    return OIT_ERR_CUDA;
  }
  const ncclResult_t r = api->AllGather(gathered + (size_t)g->rank * count, gathered, count, ncclUint32, g->comm, s);
  if(r != ncclSuccess)
  {
    err = std::string("ncclAllGather: ") + (api->GetErrorString ? api->GetErrorString(r) : "error");
    return OIT_ERR_CUDA;
  }
  const int    quads = W / 4;  // W is required to be a multiple of 4 for the split-frame mode
  const size_t total = (size_t)quads * H;
  const int    grid  = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  k_interleave_rows<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(gathered), reinterpret_cast<uint4*>(frame), quads, H, stripRows,
                                          g->world, padRows, stats);
  return 1;
}

}  // namespace oit
