// oit_peer.cu -- split frame over NVLink / NVSwitch peer memory.
//
// Every band (one process per GPU) owns a whole-frame buffer; the buffers are mapped into each other's address space
// with CUDA IPC, and the kernel that resolves a tile stores its pixels into ALL of them (oit_fused.cuh
// fusedResolveTile; k_scatter_rows for the staged path), so the exchange rides on the stores of the frame kernel
// itself instead of a collective after it.  Two flag rounds per frame keep the bands in step:
//
//   READY  "I am done reading frame n-1 of my buffer" -- signalled as the first node of frame n, awaited right before the
//          transparent pass (the geometry stage and the opaque pass absorb the skew between the bands);
//   DONE   "my strips of frame n are in your buffer"   -- signalled after the frame kernel (a kernel boundary orders its
//          peer stores before the flag store), awaited as the last node of the frame.
//
// The flags are frame sequence numbers kept in device memory, so the same captured graph is valid for every frame.
// A band that waits longer than PEER_TIMEOUT_NS gives up and raises STAT_PEER_TIMEOUT (oit_render returns an error)
// rather than hanging the GPU.
#include <cstring>
#include <string>

#include "oit_internal.h"

namespace oit {

constexpr unsigned long long PEER_TIMEOUT_NS = 20ull * 1000ull * 1000ull * 1000ull;

struct PeerState
{
  int        rank = 0, world = 1;
  size_t     frameBytes = 0;
  uint8_t*   chunk      = nullptr;  // [frame][flags], one cudaMalloc so that one IPC handle covers both
  void*      mapped[PEER_MAX]{};    // the other bands' chunks
  bool       open = false;
  PeerTable* table = nullptr;  // device copy
};

static size_t flagsOffset(size_t frameBytes) { return (frameBytes + 255) & ~(size_t)255; }

PeerState* peerCreate(int rank, int world, size_t frameBytes, void* handle64, std::string& err)
{
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  if(world < 1 || world > PEER_MAX || rank < 0 || rank >= world)
  {
    err = "peer exchange: at most 16 bands";
    return nullptr;
  }
  PeerState* ps  = new PeerState();
  ps->rank       = rank;
  ps->world      = world;
  ps->frameBytes = frameBytes;
  const size_t total = flagsOffset(frameBytes) + PEER_FLAG_WORDS * sizeof(uint32_t);
  cudaError_t  e     = cudaMalloc(&ps->chunk, total);
  if(e == cudaSuccess)
    e = cudaMemset(ps->chunk, 0, total);
  if(e == cudaSuccess)
    e = cudaMalloc(&ps->table, sizeof(PeerTable));
  if(e == cudaSuccess)
    e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if(e == cudaSuccess)
    e = cudaIpcGetMemHandle(&h, ps->chunk);
  if(e != cudaSuccess)
  {
    err = std::string("peer exchange: ") + cudaGetErrorString(e);
    cudaGetLastError();
    peerDestroy(ps);
    return nullptr;
  }
  memcpy(handle64, &h, 64);
  return ps;
}

int peerOpen(PeerState* ps, const void* handles, std::string& err)
{
  PeerTable t{};
  for(int b = 0; b < ps->world; b++)
  {
    uint8_t* base = ps->chunk;
    if(b != ps->rank)
    {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const uint8_t*)handles + (size_t)b * 64, 64);
      void*             p = nullptr;
      const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
      if(e != cudaSuccess)
      {
        err = std::string("cudaIpcOpenMemHandle (band ") + std::to_string(b) + "): " + cudaGetErrorString(e);
        cudaGetLastError();
        peerClose(ps);
        return OIT_ERR_UNSUPPORTED;
      }
      ps->mapped[b] = p;
      base          = (uint8_t*)p;
    }
    t.frame[b] = (uint32_t*)base;
    t.flags[b] = (uint32_t*)(base + flagsOffset(ps->frameBytes));
  }
  const cudaError_t e = cudaMemcpy(ps->table, &t, sizeof(t), cudaMemcpyHostToDevice);
  if(e != cudaSuccess)
  {
    err = std::string("peer exchange: ") + cudaGetErrorString(e);
    peerClose(ps);
    return OIT_ERR_CUDA;
  }
  ps->open = true;
  return OIT_OK;
}

void peerClose(PeerState* ps)
{
  if(!ps)
    return;
  for(int b = 0; b < PEER_MAX; b++)
    if(ps->mapped[b])
    {
      cudaIpcCloseMemHandle(ps->mapped[b]);
      ps->mapped[b] = nullptr;
    }
  ps->open = false;
}

void peerDestroy(PeerState* ps)
{
  if(!ps)
    return;
  peerClose(ps);
  cudaFree(ps->chunk);
  cudaFree(ps->table);
  delete ps;
}

uint32_t*        peerFrame(PeerState* ps) { return (uint32_t*)ps->chunk; }
const PeerTable* peerTable(PeerState* ps) { return ps->table; }

__device__ __forceinline__ uint32_t ldAcquireSys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void stReleaseSys(uint32_t* p, uint32_t v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globalTimerNs()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// one thread per band: flags[phase + rank] of band b = the number of the frame being rendered
// With DONE goes the band's overflow flag of this frame (a pair / clip buffer was too small: the frame will be rendered
// again), so that after the DONE round every band knows whether ANY band has to repeat the frame.
__global__ void __launch_bounds__(32) k_peer_signal(const PeerTable* __restrict__ t, int world, int rank, int phase,
                                                    const unsigned long long* __restrict__ stats)
{
  const int b = threadIdx.x;
  if(b >= world)
    return;
  const uint32_t seq = t->flags[rank][PEER_FLAG_SEQ] + 1u;
  if(phase == PEER_FLAG_DONE)
    t->flags[b][PEER_FLAG_OVF + rank] = stats[STAT_OVERFLOW] != 0ull ? seq : 0u;
  __threadfence_system();
  stReleaseSys(t->flags[b] + phase + rank, seq);
}

// one thread per band: spins until band b's flag in THIS band's buffer reaches the frame being rendered
__global__ void __launch_bounds__(32) k_peer_wait(const PeerTable* __restrict__ t, int world, int rank, int phase, unsigned long long* stats,
                                                   uint32_t* zeroWord)
{
  if(zeroWord && threadIdx.x == 0)
    *zeroWord = 0u;  // the tail of the pusher queue, reset before the frame kernel starts
  const int      b     = threadIdx.x;
  uint32_t*      local = t->flags[rank];
  const uint32_t seq   = local[PEER_FLAG_SEQ] + 1u;
  const unsigned long long tStart = globalTimerNs();
  if(b < world)
  {
    const unsigned long long t0 = tStart;
    while((int32_t)(ldAcquireSys(local + phase + b) - seq) < 0)
    {
      if(globalTimerNs() - t0 > PEER_TIMEOUT_NS)
      {
        atomicAdd(stats + STAT_PEER_TIMEOUT, 1ull);
        break;
      }
      __nanosleep(200);
    }
  }
  __syncwarp();
  __threadfence_system();
  if(b == 0)
    stats[STAT_WAIT_NS] += globalTimerNs() - tStart;  // (after the __syncwarp: the slowest band's flag has arrived)
  if(phase == PEER_FLAG_DONE)
  {
    // the overflow flags travelled with DONE (written before the flag's release store, read after its acquire load)
    const bool ovf = b < world && *reinterpret_cast<volatile uint32_t*>(local + PEER_FLAG_OVF + b) == seq;
    const bool any = __any_sync(0xffffffffu, ovf);
    if(b == 0)
    {
      stats[STAT_OVERFLOW_ANY] = any ? 1ull : 0ull;
      local[PEER_FLAG_SEQ]     = seq;
    }
  }
}

int peerSignal(PeerState* ps, int phase, const unsigned long long* stats, cudaStream_t s)
{
  k_peer_signal<<<1, 32, 0, s>>>(ps->table, ps->world, ps->rank, phase, stats);
  return 1;
}

int peerWait(PeerState* ps, int phase, unsigned long long* stats, cudaStream_t s, uint32_t* zeroWord)
{
  k_peer_wait<<<1, 32, 0, s>>>(ps->table, ps->world, ps->rank, phase, stats, zeroWord);
  return 1;
}

// staged path: copies this band's resolved rows (fin, [localRows][W]) into every band's frame at their global rows
__global__ void __launch_bounds__(256) k_scatter_rows(const PeerTable* __restrict__ t, const uint4* __restrict__ fin, int quadsPerRow, int localRows,
                                                      int stripRows, int world, int rank)
{
  const size_t total = (size_t)quadsPerRow * localRows;
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    const int    y = (int)(i / quadsPerRow), q = (int)(i - (size_t)y * quadsPerRow);
    const int    strip = y / stripRows;
    const int    gy    = (strip * world + rank) * stripRows + (y - strip * stripRows);
    const uint4  v     = fin[i];
    const size_t o     = (size_t)gy * quadsPerRow + q;
    for(int b = 0; b < world; b++)
      reinterpret_cast<uint4*>(t->frame[b])[o] = v;
  }
}

int peerScatterRows(PeerState* ps, const uint32_t* fin, int W, int localRows, int stripRows, cudaStream_t s)
{
  const int    quads = W / 4;
  const size_t total = (size_t)quads * localRows;
  if(total == 0)
    return 0;
  const int grid = (int)((total + 255) / 256 > 148 * 8 ? 148 * 8 : (total + 255) / 256);
  k_scatter_rows<<<grid, 256, 0, s>>>(ps->table, reinterpret_cast<const uint4*>(fin), quads, localRows, stripRows, ps->world, ps->rank);
  return 1;
}

}  // namespace oit
