// oit_peer.cu -- split frame over NVLink / NVSwitch peer memory.
//
// Every band (one process per GPU) owns TWO whole-frame buffers; they are mapped into each other's address space with CUDA
// IPC, and the kernel that resolves a tile stores its pixels into ALL bands' buffers (the pusher CTAs of oit_raster_ll.cu,
// fusedStorePixel of the other frame kernels, k_scatter_rows for the staged path): the exchange rides on the frame kernel
// itself instead of a collective after it.  Frame n goes to buffer n & 1 everywhere, and the bands agree on frame boundaries
// ONE FRAME LATE, so that in a stream of frames no band ever idles at a barrier for the slowest one:
//
//   k_peer_frame_begin (first node of the raster half of frame n)
//        signals READY = n + 1: "frame n + 1 may be written into my buffer (n + 1) & 1" -- that buffer held frame n - 1, and
//        what the host does with frame n - 1 was enqueued before it asked for frame n;
//        waits for READY >= n of every band (sent one frame ago, so normally there already).
//   the frame kernel(s) store into buffer n & 1 of every band.
//   k_peer_frame_end (last node)
//        signals DONE = n: "my strips of frame n are in your buffer" (the kernel boundary + a system fence order the frame
//        kernel's peer stores before the flag), together with this band's overflow flag of frame n;
//        waits for DONE >= n - 1 of every band (the frame before: normally there already) and mirrors the statistics.
//   k_peer_frame_flush (when the host wants the latest frame n: oit_synchronize, oit_download, oit_get_stats, ...)
//        waits for DONE >= n and collects the bands' overflow flags of frame n.
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
  uint8_t*   chunk      = nullptr;  // [frame 0][frame 1][flags], one cudaMalloc so that one IPC handle covers all
  void*      mapped[PEER_MAX]{};    // the other bands' chunks
  bool       open = false;
  PeerTable* table = nullptr;  // device copy
};

static size_t frameStride(size_t frameBytes) { return (frameBytes + 255) & ~(size_t)255; }
static size_t flagsOffset(size_t frameBytes) { return 2 * frameStride(frameBytes); }

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
    t.frame[0][b] = (uint32_t*)base;
    t.frame[1][b] = (uint32_t*)(base + frameStride(ps->frameBytes));
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

uint32_t*        peerFrame(PeerState* ps, unsigned which) { return (uint32_t*)(ps->chunk + (which & 1u) * frameStride(ps->frameBytes)); }
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

// spins until band b's flag `slot + b` in THIS band's page reaches `want`; returns false on time-out
__device__ __forceinline__ bool waitFlag(const uint32_t* local, int slot, int b, uint32_t want, unsigned long long* stats)
{
  const unsigned long long t0 = globalTimerNs();
  while((int32_t)(ldAcquireSys(local + slot + b) - want) < 0)
  {
    if(globalTimerNs() - t0 > PEER_TIMEOUT_NS)
    {
      atomicAdd(stats + STAT_PEER_TIMEOUT, 1ull);
      return false;
    }
    __nanosleep(200);
  }
  return true;
}

// one thread per band.  zeroA / zeroB: words the frame kernel expects cleared (linked-list counter, tail of the pusher queue)
__global__ void __launch_bounds__(32) k_peer_frame_begin(const PeerTable* __restrict__ t, int world, int rank, unsigned long long* stats,
                                                         uint32_t* zeroA, uint32_t* zeroB)
{
  const int      b     = threadIdx.x;
  uint32_t*      local = t->flags[rank];
  const uint32_t n     = local[PEER_FLAG_SEQ] + 1u;  // the frame that starts
  if(b == 0)
  {
    if(zeroA)
      *zeroA = 0u;
    if(zeroB)
      *zeroB = 0u;
  }
  const unsigned long long tStart = globalTimerNs();
  if(b < world)
  {
    stReleaseSys(t->flags[b] + PEER_FLAG_READY + rank, n + 1u);
    waitFlag(local, PEER_FLAG_READY, b, n, stats);
  }
  __syncwarp();
  if(b == 0)
    stats[STAT_WAIT_NS] += globalTimerNs() - tStart;
}

// one thread per band.  mirror: the pinned host copy of the statistics (+ the pair counts of the two draws behind them)
__global__ void __launch_bounds__(32) k_peer_frame_end(const PeerTable* __restrict__ t, int world, int rank, unsigned long long* stats,
                                                       unsigned long long* mirror, int mirrorWords, const uint32_t* __restrict__ pairInfoA,
                                                       const uint32_t* __restrict__ pairInfoB)
{
  const int      b     = threadIdx.x;
  uint32_t*      local = t->flags[rank];
  const uint32_t n     = local[PEER_FLAG_SEQ] + 1u;  // the frame that ends
  __threadfence_system();                             // the frame kernel's peer stores, made visible before the flag
  if(b < world)
  {
    t->flags[b][PEER_FLAG_OVF + (n & 1u) * PEER_MAX + rank] = stats[STAT_OVERFLOW] != 0ull ? n : 0u;
    __threadfence_system();
    stReleaseSys(t->flags[b] + PEER_FLAG_DONE + rank, n);
  }
  const unsigned long long tStart = globalTimerNs();
  if(b < world && n > 1u)
    waitFlag(local, PEER_FLAG_DONE, b, n - 1u, stats);
  __syncwarp();
  if(b == 0)
  {
    stats[STAT_WAIT_NS] += globalTimerNs() - tStart;
    local[PEER_FLAG_SEQ] = n;
  }
  __syncwarp();
  // the statistics of the frame, mirrored into pinned host memory (what two small copy nodes used to do)
  for(int i = b; i < mirrorWords; i += 32)
    mirror[i] = stats[i];
  uint32_t* m32 = reinterpret_cast<uint32_t*>(mirror + mirrorWords);
  if(b < 4)
  {
    m32[b]     = pairInfoA ? pairInfoA[b] : 0u;
    m32[4 + b] = pairInfoB ? pairInfoB[b] : 0u;
  }
  __threadfence_system();
}

// completes the latest frame (n = SEQ) for the host: every band's strips have arrived, and whether ANY band overflowed
__global__ void __launch_bounds__(32) k_peer_frame_flush(const PeerTable* __restrict__ t, int world, int rank, unsigned long long* stats,
                                                         unsigned long long* mirror)
{
  const int      b     = threadIdx.x;
  uint32_t*      local = t->flags[rank];
  const uint32_t n     = local[PEER_FLAG_SEQ];
  bool           ovf   = false;
  if(b < world && n > 0u)
  {
    waitFlag(local, PEER_FLAG_DONE, b, n, stats);
    // the overflow flags travelled with DONE (written before the flag's release store, read after its acquire load)
    ovf = *reinterpret_cast<volatile uint32_t*>(local + PEER_FLAG_OVF + (n & 1u) * PEER_MAX + b) == n;
  }
  const bool any = __any_sync(0xffffffffu, ovf);
  if(b == 0)
  {
    stats[STAT_OVERFLOW_ANY]  = any ? 1ull : 0ull;
    mirror[STAT_OVERFLOW_ANY] = any ? 1ull : 0ull;
    mirror[STAT_PEER_TIMEOUT] = stats[STAT_PEER_TIMEOUT];
    __threadfence_system();
  }
}

int peerFrameBegin(PeerState* ps, unsigned long long* stats, uint32_t* zeroA, uint32_t* zeroB, cudaStream_t s)
{
  k_peer_frame_begin<<<1, 32, 0, s>>>(ps->table, ps->world, ps->rank, stats, zeroA, zeroB);
  return 1;
}
int peerFrameEnd(PeerState* ps, unsigned long long* stats, unsigned long long* hostMirror, int mirrorWords, const uint32_t* pairInfoA,
                 const uint32_t* pairInfoB, cudaStream_t s)
{
  k_peer_frame_end<<<1, 32, 0, s>>>(ps->table, ps->world, ps->rank, stats, hostMirror, mirrorWords, pairInfoA, pairInfoB);
  return 1;
}
int peerFrameFlush(PeerState* ps, unsigned long long* stats, unsigned long long* hostMirror, cudaStream_t s)
{
  k_peer_frame_flush<<<1, 32, 0, s>>>(ps->table, ps->world, ps->rank, stats, hostMirror);
  return 1;
}

// staged path: copies this band's resolved rows (fin, [localRows][W]) into every band's frame at their global rows
__global__ void __launch_bounds__(256) k_scatter_rows(const PeerTable* __restrict__ t, const uint4* __restrict__ fin, int quadsPerRow, int localRows,
                                                      int stripRows, int world, int rank)
{
  const size_t     total  = (size_t)quadsPerRow * localRows;
  uint32_t* const* frames = peerFramesOfRunningFrame(t, rank);
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    const int    y = (int)(i / quadsPerRow), q = (int)(i - (size_t)y * quadsPerRow);
    const int    strip = y / stripRows;
    const int    gy    = (strip * world + rank) * stripRows + (y - strip * stripRows);
    const uint4  v     = fin[i];
    const size_t o     = (size_t)gy * quadsPerRow + q;
    for(int b = 0; b < world; b++)
      reinterpret_cast<uint4*>(frames[b])[o] = v;
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
