// oit_raster_ll.cu -- the Linked List colour pass (K4, oitLinkedList.frag.glsl:51-85) as an ORDER-FREE tile kernel, fused
// with the tile's composite (K5, oitLinkedList.frag.glsl:106-172) and resolve when oit_render asks for it.
//
// What the list needs from "primitive order" is only the ORDER OF THE LINKS: fragment k of a pixel points to fragment k-1.
// That order is known before anything is shaded, so nothing has to be serialised:
//
//   A  coverage   each thread tests ITEMS_PER_THREAD (triangle, pixel) candidates (oit_raster_common.cuh).  A covered
//                 fragment sets bit `triangle slot` in its pixel's 128-bit set (one chunk = 128 staged triangles, so the
//                 bit index IS the primitive order) and is appended to an unordered compact list.
//   B  allocate   thread = pixel: popcount of the set = fragments of the pixel in this batch; a warp scan gives every pixel a
//                 contiguous range of the nodes its warp reserves with ONE atomicAdd on the global counter (the reference
//                 issues one per fragment, oitLinkedList.frag.glsl:55), and the pixel's head moves to the last of them.
//   C  shade      dense over the compact list, any order: rank of the fragment among its pixel's = popcount of the lower
//                 bits; node = pixel's first node + rank; next = node - 1 (rank > 0) or the pixel's previous head.  Shade,
//                 pack, ONE 128-bit store.  No tickets, no layers, no per-fragment atomics, no barrier per layer.
//
// Two CTA barriers per batch of 2048 candidates (two coverage rounds; after A, after B; sets and lists are double-buffered
// so that C of one batch overlaps A of the next).  The list heads of the tile live in shared memory for the whole pass (written to imgAux
// once at the end); the lists are identical to the ones the sequential schedule builds (same nodes per pixel in the same
// link order; node NUMBERS differ, like between any two runs of the reference).
//
// Fused frame: the composite hands the tile's pixels to the threads in the order of their list lengths (counting sort), so
// that the lanes of a warp walk, sort and blend lists of similar length.
//
// Pool overflow (node index >= capacity, oitLinkedList.frag.glsl:82): the overflowing fragments are tail-blended by the
// ROP in primitive order.  The batch then takes a slower path: the fragments are permuted into pixel-major order, shaded
// densely in rounds of 256 into a small shared-memory queue, and each pixel's owner thread blends its fragments in order.
#include "oit_raster_common.cuh"

namespace oit {

constexpr int LL_CHUNK     = 128;  // triangles staged per chunk = bits of a pixel's per-batch triangle set
constexpr int LL_IPT       = ITEMS_PER_THREAD;
constexpr int LL_ROUND     = RASTER_THREADS * LL_IPT;  // candidates tested per coverage round
// Per sample count: coverage rounds per batch (a batch = ROUNDS * 1024 candidates between two allocation phases; its lists
// take ROUNDS * 8 KB of shared memory) and resident CTAs per SM (5: 48 registers; 4: 64 registers, no spills).  Measured on
// B200: the 8-sample instance runs best with longer batches at 4 CTAs, the others with the short batch at 5.
#ifndef OIT_LL_ROUNDS_S8
#define OIT_LL_ROUNDS_S8 2
#endif
#ifndef OIT_LL_MINB_S8
#define OIT_LL_MINB_S8 4
#endif
#ifndef OIT_LL_ROUNDS_S4
#define OIT_LL_ROUNDS_S4 2
#endif
#ifndef OIT_LL_MINB_S4
#define OIT_LL_MINB_S4 5
#endif
#ifndef OIT_LL_ROUNDS_S1
#define OIT_LL_ROUNDS_S1 2
#endif
#ifndef OIT_LL_MINB_S1
#define OIT_LL_MINB_S1 5
#endif
__host__ __device__ constexpr int llRounds(int S) { return S == 8 ? OIT_LL_ROUNDS_S8 : (S == 4 ? OIT_LL_ROUNDS_S4 : OIT_LL_ROUNDS_S1); }
__host__ __device__ constexpr int llMinBlocks(int S) { return S == 8 ? OIT_LL_MINB_S8 : (S == 4 ? OIT_LL_MINB_S4 : OIT_LL_MINB_S1); }
#ifndef OIT_LL_KNOCKOUT
#define OIT_LL_KNOCKOUT 0  // timing experiments only (tools/build_variants.sh): 1 = no composite, 2 = no shading, 4 = no node store
#endif

// ---- split frame with pusher CTAs -------------------------------------------------------------------------------------------
// The tile CTAs write their resolved pixels to `fin` and append the tile to a queue (completion order); the first
// p.pushers CTAs of the grid -- dispatched first, so resident for the whole kernel -- take queue slots round robin, one warp
// per tile, and copy the tile's 16 rows x 64 bytes into EVERY band's whole-frame buffer with 128-bit stores over NVLink.
// Compute and exchange stay one kernel, but the remote stores (7 x 4 bytes per pixel at 8 bands) are issued by a few
// warps whose only job that is: NVLink back-pressure stalls them instead of the load/store units of the rasterising SMs
// (per-pixel stores from the tile CTAs cost +28 % of the colour pass at 8 bands), and the transfers are 64-byte segments.
__device__ __forceinline__ uint32_t ldAcquireGpu(const uint32_t* a)
{
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a) : "memory");
  return v;
}
__device__ __forceinline__ void stReleaseGpu(uint32_t* a, uint32_t v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(a), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pushTimerNs()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// a tile CTA is done with its tile: every thread's stores to `fin` are ordered before the queue entry
__device__ __forceinline__ void publishTile(const FrameParams& p, uint32_t tile)
{
  if(p.pushers == 0)
    return;
  __threadfence();
  __syncthreads();
  if(threadIdx.x == 0)
  {
    const uint32_t numTiles = (uint32_t)(p.tilesX * p.tileRowsLocal);
    const uint32_t slot     = atomicAdd(p.pushQueue + numTiles, 1u);
    stReleaseGpu(p.pushQueue + slot, tile + 1u);
  }
}

__device__ __noinline__ void pushTiles(const FrameParams p)
{
  const int      lane = threadIdx.x & 31, warpsPerCta = RASTER_THREADS / 32;
  const uint32_t numTiles = (uint32_t)(p.tilesX * p.tileRowsLocal);
  const uint32_t nWarps = (uint32_t)p.pushers * warpsPerCta, myWarp = blockIdx.x * warpsPerCta + (threadIdx.x >> 5);
  const int      stripRows = p.stripTileRows * TILE_H;
  const uint4*   fin4 = reinterpret_cast<const uint4*>(p.fin);
  uint32_t* const* frames = peerFramesOfRunningFrame(p.peers, p.bandIndex);
  const int      quadsPerRow = p.W / 4;
  for(uint32_t slot = myWarp; slot < numTiles; slot += nWarps)
  {
    uint32_t v = 0;
    if(lane == 0)
    {
      const unsigned long long t0 = pushTimerNs();
      while((v = ldAcquireGpu(p.pushQueue + slot)) == 0u)
      {
        if(pushTimerNs() - t0 > 2000000000ull)
        {
          atomicAdd(p.stats + STAT_PEER_TIMEOUT, 1ull);  // (cannot happen: every tile CTA publishes; reported, not hung)
          break;
        }
        __nanosleep(64);
      }
      p.pushQueue[slot] = 0u;  // ready for the next frame
    }
    v = __shfl_sync(0xffffffffu, v, 0);
    if(v == 0u)
      return;
    const uint32_t tile = v - 1u;
    const int      rl = (int)(tile / p.tilesX), tx = (int)(tile - rl * p.tilesX);
#pragma unroll
    for(int q = lane; q < TILE_H * (TILE_W / 4); q += 32)
    {
      const int row = q / (TILE_W / 4), seg = q - row * (TILE_W / 4);
      const int yl = rl * TILE_H + row, xq = tx * (TILE_W / 4) + seg;
      if(yl < p.localH && xq < quadsPerRow)
      {
        const uint4  px    = __ldcg(fin4 + (size_t)yl * quadsPerRow + xq);
        const int    strip = yl / stripRows;
        const size_t o     = (size_t)((strip * p.bandCount + p.bandIndex) * stripRows + (yl - strip * stripRows)) * quadsPerRow + xq;
        for(int b = 0; b < p.bandCount; b++)
          reinterpret_cast<uint4*>(frames[b])[o] = px;
      }
    }
  }
}

template <int S>
__global__ void __launch_bounds__(RASTER_THREADS, llMinBlocks(S)) k_raster_ll(const FrameParams p)
{
  if(p.pushers && blockIdx.x < (unsigned)p.pushers)
  {
    pushTiles(p);
    return;
  }
  constexpr int LL_BATCH = LL_ROUND * llRounds(S);
  static_assert(RASTER_THREADS == TILE_PIX, "phase B maps one thread to one pixel of the tile");
  // per-chunk / per-batch structures; dead once the tile's list has been walked, when the fused composite reuses the space
  constexpr size_t SLOT_BYTES = sizeof(TriSlot) * LL_CHUNK;
  constexpr size_t SET_BYTES  = sizeof(uint4) * TILE_PIX * 2;        // per-pixel triangle sets of two consecutive batches
  constexpr size_t LIST_BYTES = sizeof(uint32_t) * LL_BATCH * 2;      // compact fragment lists of two consecutive batches
  constexpr size_t WORK_BYTES = SLOT_BYTES + SET_BYTES + LIST_BYTES;
  constexpr size_t SCRATCH_BYTES = WORK_BYTES > sizeof(FusedArrays) ? WORK_BYTES : sizeof(FusedArrays);
  static_assert(SLOT_BYTES % 16 == 0, "alignment of the sets");
  __shared__ __align__(16) unsigned char scratch[SCRATCH_BYTES];
  __shared__ SrgbTables tabs;
  __shared__ uint32_t   itemStart[LL_CHUNK + 1];
  __shared__ uint32_t   headSm[TILE_PIX];    // list head of every pixel of the tile
  __shared__ uint32_t   prevHead[TILE_PIX];  // ... before the current batch
  __shared__ uint32_t   pixOff[TILE_PIX];    // first node (relative to the batch's base) of the pixel's fragments
  __shared__ uint32_t   pixPre[TILE_PIX];    // fragments in set words 0..w-1, one byte per word w
  __shared__ uint32_t   sCount[2];
  __shared__ uint16_t   pixRel[TILE_PIX];    // overflow path: the pixel's first position in the batch's pixel-major order
  __shared__ uint32_t   scanSm[33];
  __shared__ uint8_t    tailMask[RASTER_THREADS];
  extern __shared__ __align__(16) unsigned char dynSmem[];  // fused frame: the tile's colour samples
  TriSlot*  slots   = reinterpret_cast<TriSlot*>(scratch);
  uint4*    pixSet  = reinterpret_cast<uint4*>(scratch + SLOT_BYTES);
  uint32_t* lists   = reinterpret_cast<uint32_t*>(scratch + SLOT_BYTES + SET_BYTES);

  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = p.tileOrder[blockIdx.x - p.pushers];  // launch order: longest lists first
  const uint32_t listBegin = p.tileStart[tile], listEnd = p.tileStart[tile + 1];
  const bool     fused = p.fused != 0;
  if(listBegin == listEnd && !fused)
    return;
  const int rl = tile / p.tilesX, tx = tile - rl * p.tilesX;
  const int R  = tileRowToGlobal(rl, p.stripTileRows, p.bandCount, p.bandIndex);
  const int tileX0 = tx * TILE_W, tileY0 = R * TILE_H;  // global pixel origin of the tile
  const int yLocal0 = rl * TILE_H;                      // the same row inside this band's buffers
  // the pixel this thread owns in phase B
  const int    ownX = tileX0 + (tid & (TILE_W - 1)), ownLy = tid >> TILE_SHIFT;
  const bool   ownValid = ownX < p.W && tileY0 + ownLy < p.H;
  const size_t ownPix   = (size_t)(yLocal0 + ownLy) * p.W + ownX;

  uint32_t* tileColorSm = reinterpret_cast<uint32_t*>(dynSmem);
  uint32_t* tileColor   = fused ? tileColorSm : nullptr;
  const bool emptyTile  = listBegin == listEnd;
  if(fused)
  {
    if(p.depth == nullptr)
    {
      // no opaque pass: the tile's colour samples start as the clear colour -- and stay that if nothing is drawn into it
      if(emptyTile)
      {
        fusedClearTile(p, tileX0, yLocal0, tid);
        if(ownValid)
          p.aux[ownPix] = 0u;  // (the fused frame does not clear imgAux beforehand)
        publishTile(p, tile);
        return;
      }
      const uint4 cc = make_uint4(p.clearColor, p.clearColor, p.clearColor, p.clearColor);
      for(int i = tid; i < TILE_PIX * S / 4; i += RASTER_THREADS)
        reinterpret_cast<uint4*>(tileColorSm)[i] = cc;
    }
    else
    {
      // ... or as what the opaque pass left in m_colorImage
      for(int i = tid; i < TILE_PIX * S; i += RASTER_THREADS)
      {
        const int pl = i / S, gx = tileX0 + (pl & (TILE_W - 1)), ly = pl >> TILE_SHIFT;
        uint32_t  v  = p.clearColor;
        if(gx < p.W && tileY0 + ly < p.H)
          v = p.color[((size_t)(yLocal0 + ly) * p.W + gx) * S + (i - pl * S)];
        tileColorSm[i] = v;
      }
    }
  }
  loadTables(tabs, p.tables);
  headSm[tid] = (ownValid && !emptyTile && !fused) ? p.aux[ownPix] : 0u;  // fused frame: every list starts empty
  if(!emptyTile)
  {
    pixSet[tid]            = make_uint4(0u, 0u, 0u, 0u);
    pixSet[TILE_PIX + tid] = make_uint4(0u, 0u, 0u, 0u);
  }
  if(tid < 2)
    sCount[tid] = 0u;
  uint32_t  nFrag = 0, nStored = 0, nTail = 0;
  uint32_t  parity = 0;
  uint32_t  ownTotal = 0;  // fragments of the pixel this thread owns, over the whole pass
  const int lo = S == 1 ? 128 : (S == 4 ? 32 : 16);  // samples sit in [lo, 256 - lo] of the pixel
  uint4*    nodes = reinterpret_cast<uint4*>(p.abuf);
  __syncthreads();

  // shading of one fragment record at the pixel centre (coverage shading / no AA: SURVEY 8a row R) -> packed colour, depth
  auto shadeRecord = [&](uint32_t rec, Color4& rgba, float& z) {
    if(OIT_LL_KNOCKOUT & 2)
    {
      rgba = Color4{0.5f, 0.25f, 0.125f, __uint_as_float(rec)};
      z    = 0.5f;
      return;
    }
    const TriSlot& s     = slots[rec & (LL_CHUNK - 1)];
    const bool     small = (s.box >> 20) & 1u;
    const int      lx = (rec >> 8) & 15, ly = (rec >> 12) & 15;
    const int      cx = ((tileX0 + lx) << 8) + 128, cy = ((tileY0 + ly) << 8) + 128;
    const Bary     bc = makeBary(edgeFloat(s, 1, cx, cy, small), edgeFloat(s, 2, cx, cy, small), s.rarea);
    float          vz = 0.f;
    rgba              = shadeAt<false>(p, s, bc, vz);
    z                 = depthAt(s, bc);
  };

  for(uint32_t base = listBegin; base < listEnd; base += LL_CHUNK)
  {
    // ---- stage: one triangle per thread (the first LL_CHUNK threads) -------------------------------------------------
    uint32_t nItems = 0;
    if(tid < LL_CHUNK && base + tid < listEnd)
    {
      const uint32_t val = p.pairTri[base + tid];
      if(!(val & PAIR_CLIPPED))
      {
        const uint32_t i0 = p.indices[3 * (size_t)val], i1 = p.indices[3 * (size_t)val + 1], i2 = p.indices[3 * (size_t)val + 2];
        nItems            = setupSlot(p.tv[i0], p.tv[i1], p.tv[i2], i0, i1, i2, 0u, p.W, p.H, tileX0, tileY0, lo, slots[tid]);
      }
      else if(val != PAIR_SKIP)
      {
        // a piece of a near-clipped triangle: its vertices come from the frame's clip table (oit_clip.cuh, k_bin_emit)
        const uint32_t   e  = val & ~PAIR_CLIPPED;
        const ClipEntry& ce = p.clipEntries[e];
        nItems              = setupSlot(ce.v[0], ce.v[1], ce.v[2], e, 0u, 0u, SLOT_CLIPPED, p.W, p.H, tileX0, tileY0, lo, slots[tid]);
      }
      else
      {
        slots[tid].box  = 0u;
        slots[tid].rcpW = 0u;
      }
    }
    uint32_t total;
    {
      const uint32_t ex = blockExclusiveScan(nItems, scanSm, total);
      if(tid <= LL_CHUNK)
        itemStart[tid] = ex;  // thread LL_CHUNK holds the total (the threads behind the chunk contribute nothing)
    }
    __syncthreads();

    for(uint32_t k0 = 0; k0 < total; k0 += LL_BATCH)
    {
      const uint32_t par      = parity & 1u;
      uint32_t*      setWords = reinterpret_cast<uint32_t*>(pixSet + par * TILE_PIX);
      uint4*         setOther = pixSet + (par ^ 1u) * TILE_PIX;
      uint32_t*      list     = lists + par * LL_BATCH;

      // ---- A: coverage of LL_IPT consecutive candidates of one triangle; compact list of the covered ones --------------------
      if(tid == 0)
        sCount[par ^ 1u] = 0u;
#pragma unroll 1
      for(int round = 0; round < llRounds(S); round++)
      {
        const uint32_t k = k0 + round * LL_ROUND + tid * LL_IPT;
        if(round > 0 && k0 + round * LL_ROUND >= total)
          break;
        uint32_t recs[LL_IPT];
        coverCandidates<S, LL_CHUNK>(p, slots, itemStart, k, total, tileX0, tileY0, yLocal0, setWords, recs);
        appendCovered(recs, &sCount[par], list);
      }
      __syncthreads();  // (1)

      // ---- B: thread = pixel.  Fragment count of the pixel, its node range, the new head ---------------------------------------
      // Every warp reserves the nodes of its 32 pixels with ONE atomicAdd on the global counter (the reference issues one
      // per fragment, oitLinkedList.frag.glsl:55): a pixel's nodes are contiguous, the order of the ranges is free.
      const uint32_t n = sCount[par];
      setOther[tid]    = make_uint4(0u, 0u, 0u, 0u);  // the next batch's sets (their last readers are past barrier 1)
      const uint4    m  = pixSet[par * TILE_PIX + tid];
      const uint32_t c0 = __popc(m.x), c1 = __popc(m.y), c2 = __popc(m.z), c3 = __popc(m.w);
      const uint32_t c  = c0 + c1 + c2 + c3;
      const uint32_t incl = warpInclusiveScan(c);
      uint32_t       wbase = 0;
      if(lane == 31 && incl)
        wbase = atomicAdd(p.counter, incl);
      wbase = __shfl_sync(0xffffffffu, wbase, 31);
      // the pixel's fragment of rank r gets node off + 1 + r (node 0 is the list terminator); nodes < capacity fit the pool
      const uint32_t off    = wbase + incl - c;
      const uint32_t room   = off + 1u < p.capacity ? p.capacity - (off + 1u) : 0u;
      const uint32_t stored = min(c, room);
      ownTotal += c;
      if(c)
      {
        pixOff[tid]   = off;
        pixPre[tid]   = (c0 << 8) | ((c0 + c1) << 16) | ((c0 + c1 + c2) << 24);
        prevHead[tid] = headSm[tid];
        if(stored)
          headSm[tid] = off + stored;  // the pixel's last stored fragment
      }
      const bool overflow = __syncthreads_or(stored < c) != 0;  // (2)

      // rank of a fragment among its pixel's fragments of this batch = position of its node inside the pixel's range
      auto fragRank = [&](uint32_t rec) {
        const uint32_t slot = rec & (LL_CHUNK - 1), pl = (rec >> 8) & 255u, w = slot >> 5;
        return __popc(setWords[pl * 4 + w] & ((1u << (slot & 31u)) - 1u)) + ((pixPre[pl] >> (8u * w)) & 255u);
      };

      if(!overflow)
      {
        // ---- C: shade + store, dense and in any order ------------------------------------------------------------------------
        for(uint32_t i = tid; i < n; i += RASTER_THREADS)
        {
          const uint32_t rec  = list[i];
          const uint32_t rank = fragRank(rec), pl = (rec >> 8) & 255u;
          const uint32_t node = pixOff[pl] + 1u + rank;
          const uint32_t next = rank ? node - 1u : prevHead[pl];
          Color4         rgba;
          float          z;
          shadeRecord(rec, rgba, z);
          if(!(OIT_LL_KNOCKOUT & 4) || rgba.r == 77.f)
            nodes[node] = make_uint4(packColor(tabs, rgba), __float_as_uint(z), S > 1 ? (rec >> 16) & 255u : 0u, next);
        }
        if(tid == 0)
        {
          nFrag += n;  // (the statistics are summed over the CTA at the end: one thread counts the whole batch)
          nStored += n;
        }
      }
      else
      {
        // ---- the pool runs out inside (or before) this batch --------------------------------------------------------------
        // fragments that still fit are stored as above; all are copied to their pixel-major position of the batch (into the
        // other batch's list buffer: idle until the next batch's phase A), which needs the CTA-wide scan of the counts
        uint32_t* sorted = lists + (par ^ 1u) * LL_BATCH;
        uint32_t  total;
        const uint32_t rel = blockExclusiveScan(c, scanSm, total);
        pixRel[tid]        = (uint16_t)rel;
        __syncthreads();
#pragma unroll 1
        for(uint32_t i = tid; i < n; i += RASTER_THREADS)
        {
          const uint32_t rec  = list[i];
          const uint32_t rank = fragRank(rec), pl = (rec >> 8) & 255u;
          const uint32_t node = pixOff[pl] + 1u + rank;
          const bool     fits = node < p.capacity;
          sorted[pixRel[pl] + rank] = fits ? 0u : rec;  // only the overflowing ones are looked at again
          nFrag++;
          if(fits)
          {
            const uint32_t next = rank ? node - 1u : prevHead[pl];
            Color4         rgba;
            float          z;
            shadeRecord(rec, rgba, z);
            nodes[node] = make_uint4(packColor(tabs, rgba), __float_as_uint(z), S > 1 ? (rec >> 16) & 255u : 0u, next);
            nStored++;
          }
          else if(p.tailBlend)
            nTail++;
        }
        __syncthreads();
        if(p.tailBlend)
        {
          // tail blend (oitLinkedList.frag.glsl:82-84 + BlendMode::PREMULTIPLIED): shaded densely, 256 list positions per
          // round, into a queue that borrows the next batch's (idle, zeroed) sets; blended by the pixel's owner in order
          float4*   queue = reinterpret_cast<float4*>(setOther);
          uint32_t* px    = tileColor ? tileColor + tid * S : p.color + ownPix * S;
          for(uint32_t r0 = 0; r0 < n; r0 += RASTER_THREADS)
          {
            const uint32_t i   = r0 + tid;
            const uint32_t rec = i < n ? sorted[i] : 0u;
            if(rec)
            {
              Color4 rgba;
              float  z;
              shadeRecord(rec, rgba, z);
              const Color4 pm = premultiply(rgba);
              queue[tid]      = make_float4(pm.r, pm.g, pm.b, pm.a);
              tailMask[tid]   = (uint8_t)((rec >> 16) & 255u);
            }
            __syncthreads();
            if(stored < c)
            {
              const uint32_t qa = max(rel + stored, r0), qb = min(rel + c, min(r0 + RASTER_THREADS, n));
              for(uint32_t q = qa; q < qb; q++)
              {
                const float4 v = queue[q - r0];
                const Color4 src{v.x, v.y, v.z, v.w};
                if(!isZero(src))
                  ropSamplesNonZero<S>(tabs, px, tailMask[q - r0], src);
              }
            }
            __syncthreads();
          }
          setOther[tid] = make_uint4(0u, 0u, 0u, 0u);
          __syncthreads();
        }
      }
      parity++;
    }
    __syncthreads();  // the slots are about to be replaced
  }

  // the heads go to imgAux (the staged composite, dumps and a later colour pass read them there)
  if((!emptyTile || fused) && ownValid)
    p.aux[ownPix] = headSm[tid];

  // ---- fused frame: composite + resolve of the tile while its nodes are still in L1 / L2 ------------------------------------
  if(fused)
  {
    if(!emptyTile)
    {
      FusedArrays& A = *reinterpret_cast<FusedArrays*>(scratch);
      // The composite's cost grows with the length of the pixel's list, and neighbouring pixels differ a lot (silhouettes,
      // background): the pixels are handed to the threads in the order of their fragment counts, longest first (a counting
      // sort of the tile's 256 pixels), so that the lanes of a warp walk, sort and blend lists of similar length.
      const uint32_t bin = 31u - min(ownTotal, 31u);
      if(tid < 32)
        scanSm[tid] = 0u;
      __threadfence_block();
      __syncthreads();
#ifdef OIT_EXPERIMENT_NO_CHASE
      prevHead[tid] = ownTotal;
#endif
      const uint32_t inBin = atomicAdd(&scanSm[bin], 1u);
      __syncthreads();
      if(warp == 0)
      {
        const uint32_t v = scanSm[lane];
        scanSm[lane]     = warpInclusiveScan(v) - v;
      }
      __syncthreads();
      pixOff[scanSm[bin] + inBin] = (uint32_t)tid | (headSm[tid] ? 256u : 0u);
      __syncthreads();
      const uint32_t mine = pixOff[tid];
      const uint32_t px   = mine & 255u;  // bit 8: a pixel with a list (pixels outside the frame have none)
#ifdef OIT_EXPERIMENT_NO_CHASE
      const AbufView av{p.abuf, headSm, (size_t)prevHead[px]};
#else
      const AbufView av{p.abuf, headSm, (size_t)TILE_PIX};
#endif
      if(p.supersample == 1)
      {
        // the output pixel is the pixel: composite, ROP and resolve in registers, no barrier in between
        const int gx = tileX0 + (int)(px & (TILE_W - 1)), ly = (int)(px >> TILE_SHIFT);
        if(gx < p.W && tileY0 + ly < p.H)
          fusedFinishPixel<S, OIT_LINKEDLIST>(p, tabs, A, tid, av, (size_t)px, tileColorSm + px * S, (mine & 256u) != 0 && !(OIT_LL_KNOCKOUT & 1), gx, yLocal0 + ly);
      }
      else
      {
        if(mine & 256u)
          fusedCompositePixel<S, OIT_LINKEDLIST>(p, tabs, A, tid, av, (size_t)px, 0, tileColorSm + px * S);
        __syncthreads();
        fusedResolveTile<S>(p, tabs, tileColorSm, tileX0, yLocal0, tid);
      }
    }
    else
    {
      __syncthreads();
      fusedResolveTile<S>(p, tabs, tileColorSm, tileX0, yLocal0, tid);
    }
    publishTile(p, tile);
  }

  // ---- statistics ----------------------------------------------------------------------------------------------------
  uint32_t vals[3] = {nFrag, nStored, nTail};
#pragma unroll
  for(int q = 0; q < 3; q++)
  {
    uint32_t v = vals[q];
#pragma unroll
    for(int d = 16; d; d >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, d);
    if(lane == 0 && v)
      atomicAdd(&p.stats[q == 0 ? STAT_FRAGMENTS : (q == 1 ? STAT_STORED : STAT_TAIL)], (unsigned long long)v);
  }
}

template <int S>
static void launchLL(const FrameParams& p, unsigned grid, cudaStream_t s)
{
  const size_t dyn = p.fused ? (size_t)TILE_PIX * S * sizeof(uint32_t) : 0;
  // static + dynamic shared memory can exceed 48 KB: opt in once per device (function attributes are per device, and
  // contexts of several GPUs may live in one process)
  static bool configured[64] = {};
  int         dev            = 0;
  cudaGetDevice(&dev);
  if(dev < 0 || dev >= 64 || !configured[dev])
  {
    if(cudaFuncSetAttribute(k_raster_ll<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TILE_PIX * S * sizeof(uint32_t))) != cudaSuccess)
      return;  // stays in cudaGetLastError(), which every stage entry point checks after its launches
    if(dev >= 0 && dev < 64)
      configured[dev] = true;
  }
  k_raster_ll<S><<<grid + (unsigned)p.pushers, RASTER_THREADS, dyn, s>>>(p);
}

// The linked-list colour pass without sample shading (no AA, MSAA with coverage masks, super-sampling)
int launchRasterLinkedList(const FrameParams& p, cudaStream_t s)
{
  const unsigned grid = (unsigned)(p.tilesX * p.tileRowsLocal);
  if(grid == 0)
    return 0;
  if(p.msaa == 1)
    launchLL<1>(p, grid, s);
  else if(p.msaa == 4)
    launchLL<4>(p, grid, s);
  else
    launchLL<8>(p, grid, s);
  return 1;
}

}  // namespace oit
