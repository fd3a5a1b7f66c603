// oit_raster.cu -- tile-ordered software rasteriser driving the per-technique fragment programs of oit_fragment.cuh.
//
// Replaces, for one geometry pass of the reference (vkCmdDrawIndexed of the transparent or opaque range):
//   fixed-function raster / early per-sample depth test / post-depth coverage   main.cpp:504-532, oitColorDepthDefines.glsl:36-37
//   the fragment shader invocations K1, K2, K4, K6, K7, K9, K11, K13, K15         (see oit_fragment.cuh)
//   ROP blending in primitive order                                              main.cpp:540-592
//
// Execution model: one CTA per 16x16 screen tile.  The CTA walks the tile's triangle list (primitive order) in
// chunks of RASTER_THREADS triangles:
//   stage    thread t sets up triangle t of the chunk (edge deltas, depth plane, bounding box clipped to the tile) in
//            shared memory; a CTA scan of the (padded) box areas gives every (triangle, pixel) candidate an item index.
//   coverage each thread tests ITEMS_PER_THREAD consecutive items of ONE triangle against the S sample positions (int32
//            edge functions for triangles up to 64 px, int64 otherwise; top-left rule; early per-sample depth test).
//   tickets  every covered fragment sets bit `thread` in a per-pixel bit set in shared memory; after one barrier its
//            ticket = popcount of the lower bits = the number of EARLIER fragments of the same pixel in this batch
//            (item order == thread order == primitive order).  Fragments are bucketed by ticket into layers.
//   shade    layer 0, then layer 1, ... are processed with one barrier between layers.  A layer holds at most one
//            fragment per pixel, so inside a layer all lanes run dense and unsynchronised with the pixel exclusively
//            owned, and across layers every pixel sees its fragments in PRIMITIVE ORDER -- what the hardware ROP
//            guarantees for the tail blend and what the ordered interlock asks for.
// A tile is owned by one CTA for the whole pass, so its A-buffer slice, aux words and colour samples stay in one SM's
// L1 and in L2.
//
// Fused frame (oit_render, p.fused): the same CTA then composites and resolves its tile (oit_fused.cuh).  The tile's colour
// samples live in dynamic shared memory for the whole pass (WBOIT: its RGBA16F / R16F targets instead), and for Loop64 /
// Spinlock (p.onChip) the tile's k-buffer slice + aux words do too, addressed with the reference's own index arithmetic
// (viewSize = 256).  Only the resolved BGRA8 pixels -- and, for the linked list, the nodes -- ever reach HBM.
//
// Tuning record (B200, 4K 8x MSAA linked list; profiles/README.md): 256 threads x 5 CTAs / SM (48 registers) beats 4 x 64
// registers and 128-thread CTAs; 4 items per thread beats 2 and 8; __match_any_sync bucketing beats ballots; per-pixel
// sequence counters with spin-waits instead of the per-layer barriers were slower (two block fences per fragment).
#include <cstdlib>

#include "oit_raster_common.cuh"

namespace oit {

// one fragment record (slot | lx << 8 | ly << 12 | mask << 16) with its pixel exclusively owned
template <int PASS, int S, bool SSHADE>
__device__ __forceinline__ void processFragment(FragCtx& ctx, const TriSlot& s, uint32_t rec, int tileX0, int tileY0, int yLocal0,
                                                uint32_t* tileColor)
{
  const FrameParams& p     = ctx.p;
  const bool         small = (s.box >> 20) & 1u;
  const int          lx = (rec >> 8) & 15, ly = (rec >> 12) & 15;
  const uint32_t     mask = rec >> 16;
  const int          gx = tileX0 + lx, yl = yLocal0 + ly;
  const int          ox = gx << 8, oy = (tileY0 + ly) << 8;
  // the pixel's colour samples: the shared-memory tile of the fused frame kernel, else m_colorImage
  uint32_t* colorPx = tileColor ? tileColor + (ly * TILE_W + lx) * S : p.color + ((size_t)yl * p.W + gx) * S;
  if(PASS == PASS_OPAQUE)
  {
    // opaque.frag.glsl:30-34, BlendMode::NONE with depth write (main.cpp:541-546); shaded at the pixel centre
    float        vz;
    const Bary   bc = makeBary(edgeFloat(s, 1, ox + 128, oy + 128, small), edgeFloat(s, 2, ox + 128, oy + 128, small), s.rarea);
    Color4       g  = shadeAt<false>(p, s, bc, vz);
    g.a             = 1.0f;
    const uint32_t enc = encodeDst(ctx.t, g);
    const size_t   pix = (size_t)yl * p.W + gx;
    bool           wrote = false;
#pragma unroll
    for(int sI = 0; sI < S; sI++)
      if(mask & (1u << sI))
      {
        const int   px = ox + SamplePattern<S>::x(sI), py = oy + SamplePattern<S>::y(sI);
        const Bary  b  = makeBary(edgeFloat(s, 1, px, py, small), edgeFloat(s, 2, px, py, small), s.rarea);
        const float zs = depthAt(s, b);
        // the mask was computed before this pixel's earlier fragments of the same batch ran: test again
        if(zs < p.depth[pix * S + sI])
        {
          p.depth[pix * S + sI] = zs;
          p.color[pix * S + sI] = enc;
          wrote                 = true;
        }
      }
    ctx.nOpaque += wrote ? 1u : 0u;
  }
  else if(SSHADE && PASS != PASS_WEIGHTED)
  {
    // sample shading: every covered sample is its own invocation at the sample position
#pragma unroll 1
    for(int sI = 0; sI < S; sI++)
      if(mask & (1u << sI))
      {
        const int      px = ox + SamplePattern<S>::x(sI), py = oy + SamplePattern<S>::y(sI);
        const Bary     b  = makeBary(edgeFloat(s, 1, px, py, small), edgeFloat(s, 2, px, py, small), s.rarea);
        float          vz    = 0.f;
        const uint32_t token = preInvoke<PASS>(ctx, ctx.onChip ? (size_t)(ly * TILE_W + lx) : (size_t)yl * p.W + gx, (uint32_t)sI);
        const Color4   rgba  = shadeAt<false>(p, s, b, vz);
        invoke<PASS, S>(ctx, gx, yl, (uint32_t)sI, 1u << sI, rgba, depthAt(s, b), vz, token, colorPx, ly * TILE_W + lx);
      }
  }
  else
  {
    // one invocation per pixel, varyings and gl_FragCoord.z at the pixel centre
    const Bary     bc = makeBary(edgeFloat(s, 1, ox + 128, oy + 128, small), edgeFloat(s, 2, ox + 128, oy + 128, small), s.rarea);
    float          vz    = 0.f;
    const uint32_t token = preInvoke<PASS>(ctx, ctx.onChip ? (size_t)(ly * TILE_W + lx) : (size_t)yl * p.W + gx, 0u);
    const Color4   rgba  = shadeAt<PASS == PASS_WEIGHTED>(p, s, bc, vz);
    invoke<PASS, S>(ctx, gx, yl, 0u, mask, rgba, depthAt(s, bc), vz, token, colorPx, ly * TILE_W + lx);
  }
}

// five 256-thread CTAs per SM (<= 51 registers): measured best on B200 (4 CTAs / 64 registers: +6 %, 6 CTAs spill)
#ifndef OIT_TILE_ORDER
#define OIT_TILE_ORDER 1
#endif
#ifndef OIT_MIN_BLOCKS
#define OIT_MIN_BLOCKS 5
#endif
template <int PASS, int S, bool SSHADE>
__global__ void __launch_bounds__(RASTER_THREADS, OIT_MIN_BLOCKS) k_raster(const FrameParams p)
{
  // the per-chunk structures (triangle slots, bucketed fragment records, per-pixel thread bit sets) are dead once the
  // tile's list has been walked: the fused composite reuses their shared memory for its per-pixel fragment arrays
  constexpr size_t SLOT_BYTES = sizeof(TriSlot) * RASTER_THREADS, SORTED_BYTES = sizeof(uint32_t) * BATCH_ITEMS,
                   MASK_BYTES = sizeof(uint32_t) * TILE_PIX * MASK_WORDS;
  constexpr size_t SCRATCH_BYTES = SLOT_BYTES + SORTED_BYTES + MASK_BYTES > sizeof(FusedArrays) ? SLOT_BYTES + SORTED_BYTES + MASK_BYTES : sizeof(FusedArrays);
  __shared__ __align__(16) unsigned char scratch[SCRATCH_BYTES];
  __shared__ SrgbTables tabs;
  __shared__ uint32_t   itemStart[RASTER_THREADS + 1];
  __shared__ uint32_t   layerCount[2][RASTER_THREADS];      // double buffered across batches
  extern __shared__ __align__(16) unsigned char dynSmem[];  // fused frame kernel only: see rasterDynamicSmem()
  TriSlot*  slots   = reinterpret_cast<TriSlot*>(scratch);
  uint32_t* sorted  = reinterpret_cast<uint32_t*>(scratch + SLOT_BYTES);               // fragment records bucketed by layer
  uint32_t* pixMask = reinterpret_cast<uint32_t*>(scratch + SLOT_BYTES + SORTED_BYTES);  // per pixel: which threads hold a fragment
  __shared__ uint32_t   numLayers[2];
  __shared__ uint32_t   scanSm[33];

  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = OIT_TILE_ORDER ? p.tileOrder[blockIdx.x] : blockIdx.x;  // launch order: longest lists first
  const uint32_t listBegin = p.tileStart[tile], listEnd = p.tileStart[tile + 1];
  const bool     fused = PASS != PASS_OPAQUE && PASS != PASS_LOOP_DEPTH && p.fused != 0;
  if(listBegin == listEnd && !fused)
    return;
  const int rl = tile / p.tilesX, tx = tile - rl * p.tilesX;
  const int R  = tileRowToGlobal(rl, p.stripTileRows, p.bandCount, p.bandIndex);
  const int tileX0 = tx * TILE_W, tileY0 = R * TILE_H;  // global pixel origin of the tile
  const int yLocal0 = rl * TILE_H;                      // the same row inside this band's buffers

  // dynamic shared memory of the fused frame kernel: the tile's colour samples -- or, for WBOIT, the tile's RGBA16F
  // accumulator + R16F revealage samples (WBOIT does not touch the colour target until its composite, which then uses
  // the scratch area for the colour tile)
  constexpr bool WEIGHTED     = PASS == PASS_WEIGHTED;
  uint32_t*      tileColorSm  = reinterpret_cast<uint32_t*>(WEIGHTED ? scratch : dynSmem);
  uint2*         wAccSm       = reinterpret_cast<uint2*>(dynSmem);
  uint16_t*      wRevSm       = reinterpret_cast<uint16_t*>(dynSmem + sizeof(uint2) * TILE_PIX * S);
  auto           initColorTile = [&]() {
    // the tile's colour samples start as the cleared (or opaque-drawn) m_colorImage content
    for(int i = tid; i < TILE_PIX * S; i += RASTER_THREADS)
    {
      const int pl = i / S, gx = tileX0 + (pl & (TILE_W - 1)), ly = pl >> TILE_SHIFT;
      uint32_t  v  = p.clearColor;
      if(p.depth != nullptr && gx < p.W && tileY0 + ly < p.H)
        v = p.color[((size_t)(yLocal0 + ly) * p.W + gx) * S + (i - pl * S)];
      tileColorSm[i] = v;
    }
  };
  uint32_t* tileColor = (fused && !WEIGHTED) ? tileColorSm : nullptr;
  // fused frame, nothing transparent touches this tile: only the resolve of what is there runs (same code as below)
  const bool emptyTile = listBegin == listEnd;
  if(fused)
  {
    if(!WEIGHTED || emptyTile)
      initColorTile();
    if(WEIGHTED && !emptyTile)
      for(int i = tid; i < TILE_PIX * S; i += RASTER_THREADS)
      {
        wAccSm[i] = make_uint2(0u, 0u);  // accum cleared to 0, reveal to 1.0 (oitRender.cpp:394-397)
        wRevSm[i] = 0x3C00u;
      }
  }
  loadTables(tabs, p.tables);
  if(!emptyTile)
    for(int i = tid; i < TILE_PIX * MASK_WORDS; i += RASTER_THREADS)
      pixMask[i] = 0u;
  layerCount[0][tid] = 0u;
  layerCount[1][tid] = 0u;
  if(tid < 2)
    numLayers[tid] = 0u;
  FragCtx ctx{p, tabs, 0, 0, 0, 0, (fused && WEIGHTED) ? wAccSm : nullptr, (fused && WEIGHTED) ? wRevSm : nullptr,
              p.abuf, p.aux, p.adepth, p.spin, (size_t)p.W * p.localH, false};
  if(fused && p.onChip && !WEIGHTED && !emptyTile)
  {
    // the tile's k-buffer slice + aux words in shared memory, behind the colour tile: [A-buffer][imgAux][imgDepth][imgSpin],
    // cleared like clearTransparent{Simple,Loop64,Lock} clear the global ones (oitRender.cpp:156-174,303-311,337-356)
    uint32_t*      base      = reinterpret_cast<uint32_t*>(dynSmem) + TILE_PIX * S;
    const uint32_t abufWords = onChipAbufWords(p.algorithm, p.L, p.coverage);
    ctx.abuf     = base;
    ctx.aux      = base + abufWords;
    ctx.adepth   = ctx.aux + TILE_PIX;
    ctx.spin     = ctx.adepth + TILE_PIX;
    ctx.viewSize = TILE_PIX;
    ctx.onChip   = true;
    const uint32_t abufFill = p.algorithm == OIT_LOOP64 ? 0xFFFFFFFFu : 0u;
    for(uint32_t i = tid; i < abufWords; i += RASTER_THREADS)
      base[i] = abufFill;
    for(int i = tid; i < TILE_PIX; i += RASTER_THREADS)
    {
      ctx.aux[i]    = 0u;
      ctx.adepth[i] = 0xFFFFFFFFu;
      ctx.spin[i]   = 0u;
    }
  }
  uint32_t parity = 0;
  const int lo = S == 1 ? 128 : (S == 4 ? 32 : 16);  // samples sit in [lo, 256 - lo] of the pixel
  __syncthreads();

  for(uint32_t base = listBegin; base < listEnd; base += RASTER_THREADS)
  {
    // ---- stage one triangle per thread ----------------------------------------------------------------------------
    uint32_t nItems = 0;
    if(base + tid < listEnd)
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
      itemStart[tid]    = ex;
      if(tid == 0)
        itemStart[RASTER_THREADS] = total;
    }
    __syncthreads();

    for(uint32_t k0 = 0; k0 < total; k0 += BATCH_ITEMS)
    {
      uint32_t* lcount = layerCount[parity & 1u];
      uint32_t* lnext  = layerCount[(parity & 1u) ^ 1u];
      uint32_t& nlay   = numLayers[parity & 1u];

      // ---- coverage: ITEMS_PER_THREAD consecutive items of one triangle ---------------------------------------------
      uint32_t       recs[ITEMS_PER_THREAD];
      const uint32_t k = k0 + tid * ITEMS_PER_THREAD;
#pragma unroll
      for(int j = 0; j < ITEMS_PER_THREAD; j++)
        recs[j] = 0u;
      if(k < total)
      {
        int slot = 0;
#pragma unroll
        for(int step = RASTER_THREADS / 2; step; step >>= 1)
          if(itemStart[slot + step] <= k)
            slot += step;
        const TriSlot& s      = slots[slot];
        const uint32_t bw     = ((s.box >> 8) & 15u) + 1u, bh = ((s.box >> 12) & 15u) + 1u;
        const uint32_t local0 = k - itemStart[slot];
#pragma unroll
        for(int j = 0; j < ITEMS_PER_THREAD; j++)
        {
          const uint32_t local = local0 + j;
          if(local < bw * bh)
          {
            const uint32_t row  = (local * s.rcpW) >> 16;
            const int      lx   = (int)((s.box & 15u) + (local - row * bw));
            const int      ly   = (int)(((s.box >> 4) & 15u) + row);
            const float*   dpx  = p.depth ? p.depth + ((size_t)(yLocal0 + ly) * p.W + tileX0 + lx) * S : nullptr;
            const uint32_t mask = coverageMask<S>(s, tileX0 + lx, tileY0 + ly, dpx);
            if(mask)
            {
              recs[j] = (uint32_t)slot | ((uint32_t)lx << 8) | ((uint32_t)ly << 12) | (mask << 16);
              atomicOr(&pixMask[(ly * TILE_W + lx) * MASK_WORDS + warp], 1u << lane);
            }
          }
        }
      }
      __syncthreads();

      // ---- tickets: earlier fragments of the same pixel; position inside the layer -----------------------------------
      uint32_t tick[ITEMS_PER_THREAD], pos[ITEMS_PER_THREAD];
      uint32_t maxTicket = 0;
#pragma unroll
      for(int j = 0; j < ITEMS_PER_THREAD; j++)
      {
        uint32_t t = 0xFFFFFFFFu;
        if(recs[j])
        {
          const uint32_t* m = pixMask + (((recs[j] >> 12) & 15u) * TILE_W + ((recs[j] >> 8) & 15u)) * MASK_WORDS;
          t                 = __popc(m[warp] & ((1u << lane) - 1u));
          for(int w = 0; w < warp; w++)
            t += __popc(m[w]);
          maxTicket = max(maxTicket, t + 1u);
        }
        tick[j] = t;
        // claim a slot in the layer's bucket: warp-aggregated for the common tickets 0..2, per thread beyond
        uint32_t myPos = 0;
#if OIT_TICKET_MATCH
        {
          const uint32_t peers  = __match_any_sync(0xffffffffu, t);
          const int      leader = __ffs(peers) - 1;
          uint32_t       first  = 0;
          if(lane == leader && t != 0xFFFFFFFFu)
            first = atomicAdd(&lcount[t], __popc(peers));
          myPos = __shfl_sync(0xffffffffu, first, leader) + __popc(peers & ((1u << lane) - 1u));
        }
#else
#pragma unroll
        for(uint32_t tv = 0; tv < 3; tv++)
        {
          const uint32_t votes = __ballot_sync(0xffffffffu, t == tv);
          if(votes)
          {
            const int leader = __ffs(votes) - 1;
            uint32_t  first  = 0;
            if(lane == leader)
              first = atomicAdd(&lcount[tv], __popc(votes));
            first = __shfl_sync(0xffffffffu, first, leader);
            if(t == tv)
              myPos = first + __popc(votes & ((1u << lane) - 1u));
          }
        }
        if(t != 0xFFFFFFFFu && t >= 3u)
          myPos = atomicAdd(&lcount[t], 1u);
#endif
        pos[j] = myPos;
      }
      maxTicket = __reduce_max_sync(0xffffffffu, maxTicket);
      if(lane == 0 && maxTicket)
        atomicMax(&nlay, maxTicket);
      __syncthreads();

      // ---- bucket the records by layer; reset the bit sets and the other batch's counters ----------------------------
      const uint32_t c0 = lcount[0], c1 = lcount[1], c2 = lcount[2];
#pragma unroll
      for(int j = 0; j < ITEMS_PER_THREAD; j++)
        if(recs[j])
        {
          const uint32_t t = tick[j];
          uint32_t       b = (t > 0u ? c0 : 0u) + (t > 1u ? c1 : 0u) + (t > 2u ? c2 : 0u);
          for(uint32_t l = 3; l < t; l++)
            b += lcount[l];
          sorted[b + pos[j]] = recs[j];
          pixMask[(((recs[j] >> 12) & 15u) * TILE_W + ((recs[j] >> 8) & 15u)) * MASK_WORDS + warp] = 0u;
        }
      lnext[tid] = 0u;
      if(tid == 0)
        numLayers[(parity & 1u) ^ 1u] = 0u;
      __syncthreads();

      // ---- shade + insert, layer by layer (<= 1 fragment per pixel inside a layer) --------------------------------------
      const uint32_t nl    = nlay;
      uint32_t       start = 0;
      for(uint32_t l = 0; l < nl; l++)
      {
        const uint32_t cnt = lcount[l];
        for(uint32_t i = start + tid; i < start + cnt; i += RASTER_THREADS)
        {
          const uint32_t rec = sorted[i];
          processFragment<PASS, S, SSHADE>(ctx, slots[rec & 255u], rec, tileX0, tileY0, yLocal0, tileColor);
        }
        start += cnt;
        __syncthreads();
      }
      parity++;
    }
    __syncthreads();
  }

  // ---- fused frame: composite + resolve of the tile while its A-buffer slice is still in L1 / L2 --------------------------
  if(fused)
  {
    if(!emptyTile)
    {
      FusedArrays& A = *reinterpret_cast<FusedArrays*>(scratch);
      __threadfence_block();
      __syncthreads();
      if(WEIGHTED)
      {
        initColorTile();  // into the scratch area, which the chunk structures no longer need
        __syncthreads();
      }
      for(int pl = tid; pl < TILE_PIX; pl += RASTER_THREADS)
      {
        const int gx = tileX0 + (pl & (TILE_W - 1)), ly = pl >> TILE_SHIFT;
        if(gx < p.W && tileY0 + ly < p.H)
        {
          const AbufView av{ctx.abuf, ctx.aux, ctx.viewSize};
          const size_t   pixG = (size_t)(yLocal0 + ly) * p.W + gx;
          fusedCompositePixel<S, passAlgorithm(PASS)>(p, tabs, A, pl, av, ctx.onChip ? (size_t)pl : pixG, pixG, tileColorSm + pl * S,
                                                      WEIGHTED ? wAccSm + pl * S : nullptr, WEIGHTED ? wRevSm + pl * S : nullptr);
        }
      }
    }
    __syncthreads();
    fusedResolveTile<S>(p, tabs, tileColorSm, tileX0, yLocal0, tid);
  }

  // ---- statistics ----------------------------------------------------------------------------------------------------
  uint32_t vals[4] = {ctx.nFrag, ctx.nStored, ctx.nTail, ctx.nOpaque};
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    uint32_t v = vals[q];
#pragma unroll
    for(int d = 16; d; d >>= 1)
      v += __shfl_xor_sync(0xffffffffu, v, d);
    if((tid & 31) == 0 && v)
      atomicAdd(&p.stats[q == 0 ? STAT_FRAGMENTS : (q == 1 ? STAT_STORED : (q == 2 ? STAT_TAIL : STAT_OPAQUE))], (unsigned long long)v);
  }
}

// ---- dispatch ----------------------------------------------------------------------------------------------------------
template <int PASS, int S, bool SSHADE>
static void launchKernel(const FrameParams& p, unsigned grid, cudaStream_t s)
{
  // eight 128-thread CTAs (eight tiles) per SM need ~160 KB of shared memory: ask for a large carveout once
  // dynamic shared memory of the fused frame kernel: the colour tile, or the RGBA16F + R16F WBOIT tiles (10 B / sample)
  // (+ the tile's k-buffer slice and aux words when the technique runs on chip)
  const size_t dynBytes = (size_t)TILE_PIX * S * (PASS == PASS_WEIGHTED ? 10 : 4)
                          + (p.onChip ? (size_t)onChipWords(p.algorithm, p.L, p.coverage) * 4 : 0);
  // function attributes are per device: one flag per ordinal (contexts of several GPUs may live in one process)
  static bool configured[64] = {};
  int         dev            = 0;
  cudaGetDevice(&dev);
  if(dev < 0 || dev >= 64 || !configured[dev])
  {
    cudaError_t e = cudaFuncSetAttribute(k_raster<PASS, S, SSHADE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(TILE_PIX * S * 10 + ON_CHIP_MAX_BYTES));
    if(e == cudaSuccess && OIT_SMEM_CARVEOUT >= 0)
      e = cudaFuncSetAttribute(k_raster<PASS, S, SSHADE>, cudaFuncAttributePreferredSharedMemoryCarveout, OIT_SMEM_CARVEOUT);
    if(e != cudaSuccess)
      return;  // stays in cudaGetLastError(), which every stage entry point checks after its launches
    if(dev >= 0 && dev < 64)
      configured[dev] = true;
  }
  const bool fused = p.fused && PASS != PASS_OPAQUE && PASS != PASS_LOOP_DEPTH;
  k_raster<PASS, S, SSHADE><<<grid, RASTER_THREADS, fused ? dynBytes : 0, s>>>(p);
}

template <int PASS>
static void launchPass(const FrameParams& p, cudaStream_t s)
{
  const unsigned grid = (unsigned)(p.tilesX * p.tileRowsLocal);
  const bool     ss   = p.sampleShading != 0;
  if(p.msaa == 1)
    launchKernel<PASS, 1, false>(p, grid, s);
  else if(p.msaa == 4)
  {
    if(ss)
      launchKernel<PASS, 4, true>(p, grid, s);
    else
      launchKernel<PASS, 4, false>(p, grid, s);
  }
  else
  {
    if(ss)
      launchKernel<PASS, 8, true>(p, grid, s);
    else
      launchKernel<PASS, 8, false>(p, grid, s);
  }
}

// OIT_B200_LAYERED=1: every pass through the layered kernel of this file (the round-1 design: tickets + one barrier per
// layer), kept for A/B comparisons and as a second implementation the tests run against the same checker;
// OIT_B200_LAYERED_LL=1: only the linked list.  (Read when a frame is issued or captured, not per replay.)
static bool useLayered() { return getenv("OIT_B200_LAYERED") != nullptr; }
static bool useLayeredLinkedList() { return useLayered() || getenv("OIT_B200_LAYERED_LL") != nullptr; }

// the fused linked-list frame is rendered by k_raster_ll, which starts every list empty by itself (no imgAux clear needed)
bool linkedListFrameStartsEmpty(const FrameParams& p) { return p.fused && !p.sampleShading && !useLayeredLinkedList(); }

int launchRaster(const FrameParams& p, int pass, cudaStream_t s)
{
  if(p.tilesX * p.tileRowsLocal == 0)
    return 0;
  // without sample shading the linked list has its own order-free kernel (oit_raster_ll.cu); every other pass shades
  // densely and inserts in primitive order (oit_raster_q.cu)
  if(pass == PASS_LINKEDLIST && !p.sampleShading && !useLayeredLinkedList())
    return launchRasterLinkedList(p, s);
  if(!useLayered() && !(pass == PASS_LINKEDLIST && useLayeredLinkedList()))
    return launchRasterQueued(p, pass, s);
  switch(pass)
  {
    case PASS_SIMPLE: launchPass<PASS_SIMPLE>(p, s); break;
    case PASS_LINKEDLIST: launchPass<PASS_LINKEDLIST>(p, s); break;
    case PASS_LOOP_COLOR: launchPass<PASS_LOOP_COLOR>(p, s); break;
    case PASS_LOOP64: launchPass<PASS_LOOP64>(p, s); break;
    case PASS_SPINLOCK: launchPass<PASS_SPINLOCK>(p, s); break;
    case PASS_INTERLOCK: launchPass<PASS_INTERLOCK>(p, s); break;
    case PASS_WEIGHTED: launchPass<PASS_WEIGHTED>(p, s); break;
    case PASS_LOOP_DEPTH: launchPass<PASS_LOOP_DEPTH>(p, s); break;
    case PASS_OPAQUE: launchPass<PASS_OPAQUE>(p, s); break;
    default: return 0;
  }
  return 1;
}

}  // namespace oit
