// oit_raster_q.cu -- the colour passes whose result depends on the ORDER in which a pixel sees its fragments (Simple, Loop32,
// Loop64, Spinlock, Interlock, WBOIT, the opaque pass, and the linked list under sample shading), as a tile kernel that
// shades densely and inserts in primitive order:
//
//   A  coverage   as in oit_raster_ll.cu: ITEMS_PER_THREAD candidates per thread, a compact unordered list of the covered
//                 ones, and per pixel the set of the batch's triangle slots that cover it (bit index = primitive order)
//   B  count      thread = pixel: popcount of the set, CTA scan -> where the pixel's fragments sit in pixel-major order
//   P  permute    every fragment moves to its pixel-major position (rank inside the pixel = popcount of the lower bits)
//   then, 256 queue entries at a time:
//   C  shade      dense, any order: interpolation + shading of one invocation per thread (per covered SAMPLE under sample
//                 shading) into a small shared-memory queue -- the expensive, order-free part of the fragment shader
//   D  insert     thread = pixel owner: walks ITS fragments of the round in primitive order and runs the technique's
//                 insert (oit_fragment.cuh, "owned" programs: plain read-modify-writes instead of the GLSL's atomics /
//                 spin lock / interlock) + the ROP tail blend.  What the hardware guarantees with primitive-ordered ROPs
//                 and pixel_interlock_ordered falls out of the data layout; there is no arbitration and no barrier per layer.
//
// Replaces K1, K2, K6, K7, K9, K11, K13, K15 (+ K4 under sample shading) and the fixed-function raster / early depth / ROP
// around them (main.cpp:504-592).  Fused frame (oit_render): the same CTA then composites and resolves its tile
// (oit_fused.cuh); the tile's colour samples -- for WBOIT its RGBA16F / R16F targets -- live in shared memory, and for
// Loop64 / Spinlock (p.onChip) so does the tile's k-buffer slice.
#include "oit_raster_common.cuh"

namespace oit {

constexpr int Q_CHUNK = 128;  // triangles staged per chunk = bits of a pixel's per-batch triangle set
constexpr int Q_IPT   = ITEMS_PER_THREAD;
constexpr int Q_BATCH = RASTER_THREADS * Q_IPT;
constexpr int Q_ROUND = RASTER_THREADS;  // queue entries (invocations) shaded per round
#ifndef OIT_Q_MIN_BLOCKS
#define OIT_Q_MIN_BLOCKS 4
#endif

template <int PASS, int S, bool SSHADE>
__global__ void __launch_bounds__(RASTER_THREADS, OIT_Q_MIN_BLOCKS) k_raster_q(const FrameParams p)
{
  static_assert(RASTER_THREADS == TILE_PIX, "phases B and D map one thread to one pixel of the tile");
  constexpr bool WEIGHTED  = PASS == PASS_WEIGHTED;
  constexpr bool PERSAMPLE = SSHADE && !WEIGHTED && PASS != PASS_OPAQUE;  // one invocation per covered sample
  constexpr int  ENTRIES   = PERSAMPLE ? S : 1;                           // queue entries per fragment
  constexpr int  FRAGS_PER_ROUND = Q_ROUND / ENTRIES;
  // per-chunk / per-batch structures; dead once the tile's list has been walked, when the fused composite reuses the space
  constexpr size_t SLOT_BYTES = sizeof(TriSlot) * Q_CHUNK;
  constexpr size_t SET_BYTES  = sizeof(uint4) * TILE_PIX * 2;     // per-pixel triangle sets of two consecutive batches
  constexpr size_t LIST_BYTES = sizeof(uint32_t) * Q_BATCH * 2;   // compact list + its pixel-major permutation
  constexpr size_t WORK_BYTES = SLOT_BYTES + SET_BYTES + LIST_BYTES;
  constexpr size_t SCRATCH_BYTES = WORK_BYTES > sizeof(FusedArrays) ? WORK_BYTES : sizeof(FusedArrays);
  static_assert(SLOT_BYTES % 16 == 0, "alignment of the sets");
  __shared__ __align__(16) unsigned char scratch[SCRATCH_BYTES];
  __shared__ SrgbTables tabs;
  __shared__ uint32_t   itemStart[Q_CHUNK + 1];
  __shared__ uint32_t   pixOff[TILE_PIX];  // position of the pixel's first fragment in the batch's pixel-major order
  __shared__ uint32_t   pixPre[TILE_PIX];  // fragments in set words 0..w-1, one byte per word w
  __shared__ __align__(16) uint32_t warpTot[RASTER_THREADS / 32];
  __shared__ uint32_t   sCount[2];
  __shared__ uint32_t   scanSm[33];
  // the queue of one round: what the insert needs from the shaded invocation
  __shared__ float4 qColor[Q_ROUND];  // unpremultiplied linear rgba (opaque pass: .x = the encoded BGRA8 word)
  __shared__ float  qDepth[Q_ROUND];  // gl_FragCoord.z (WBOIT: the view-space depth)
  extern __shared__ __align__(16) unsigned char dynSmem[];  // fused frame kernel only
  TriSlot*  slots  = reinterpret_cast<TriSlot*>(scratch);
  uint4*    pixSet = reinterpret_cast<uint4*>(scratch + SLOT_BYTES);
  uint32_t* lists  = reinterpret_cast<uint32_t*>(scratch + SLOT_BYTES + SET_BYTES);

  const int      tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = p.tileOrder[blockIdx.x];  // launch order: longest lists first
  const uint32_t listBegin = p.tileStart[tile], listEnd = p.tileStart[tile + 1];
  const bool     fused = PASS != PASS_OPAQUE && PASS != PASS_LOOP_DEPTH && p.fused != 0;
  if(listBegin == listEnd && !fused)
    return;
  const int rl = tile / p.tilesX, tx = tile - rl * p.tilesX;
  const int R  = tileRowToGlobal(rl, p.stripTileRows, p.bandCount, p.bandIndex);
  const int tileX0 = tx * TILE_W, tileY0 = R * TILE_H;  // global pixel origin of the tile
  const int yLocal0 = rl * TILE_H;                      // the same row inside this band's buffers
  // the pixel this thread owns in phases B and D
  const int    ownLx = tid & (TILE_W - 1), ownLy = tid >> TILE_SHIFT;
  const int    ownX  = tileX0 + ownLx;
  const bool   ownValid = ownX < p.W && tileY0 + ownLy < p.H;
  const size_t ownPix   = (size_t)(yLocal0 + ownLy) * p.W + ownX;

  // dynamic shared memory of the fused frame kernel: the tile's colour samples -- or, for WBOIT, the tile's RGBA16F
  // accumulator + R16F revealage samples (WBOIT does not touch the colour target until its composite, which then uses
  // the scratch area for the colour tile)
  uint32_t* tileColorSm  = reinterpret_cast<uint32_t*>(WEIGHTED ? scratch : dynSmem);
  uint2*    wAccSm       = reinterpret_cast<uint2*>(dynSmem);
  uint16_t* wRevSm       = reinterpret_cast<uint16_t*>(dynSmem + sizeof(uint2) * TILE_PIX * S);
  auto      initColorTile = [&]() {
    // the tile's colour samples start as the cleared (or opaque-drawn) m_colorImage content
    if(p.depth == nullptr)
    {
      const uint4 cc = make_uint4(p.clearColor, p.clearColor, p.clearColor, p.clearColor);
      for(int i = tid; i < TILE_PIX * S / 4; i += RASTER_THREADS)
        reinterpret_cast<uint4*>(tileColorSm)[i] = cc;
      return;
    }
    for(int i = tid; i < TILE_PIX * S; i += RASTER_THREADS)
    {
      const int pl = i / S, gx = tileX0 + (pl & (TILE_W - 1)), ly = pl >> TILE_SHIFT;
      uint32_t  v  = p.clearColor;
      if(gx < p.W && tileY0 + ly < p.H)
        v = p.color[((size_t)(yLocal0 + ly) * p.W + gx) * S + (i - pl * S)];
      tileColorSm[i] = v;
    }
  };
  uint32_t*  tileColor = (fused && !WEIGHTED) ? tileColorSm : nullptr;
  const bool emptyTile = listBegin == listEnd;
  if(fused)
  {
    // nothing transparent touches this tile and no opaque pass ran: every sample keeps the clear colour
    if(emptyTile && p.depth == nullptr)
    {
      fusedClearTile(p, tileX0, yLocal0, tid);
      return;
    }
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
  {
    pixSet[tid]            = make_uint4(0u, 0u, 0u, 0u);
    pixSet[TILE_PIX + tid] = make_uint4(0u, 0u, 0u, 0u);
  }
  if(tid < 2)
    sCount[tid] = 0u;
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
    ctx.aux[tid]    = 0u;
    ctx.adepth[tid] = 0xFFFFFFFFu;
    ctx.spin[tid]   = 0u;
  }
  uint32_t  parity = 0;
  const int lo = S == 1 ? 128 : (S == 4 ? 32 : 16);  // samples sit in [lo, 256 - lo] of the pixel
  // the owner's colour samples: the shared-memory tile of the fused frame kernel, else m_colorImage
  uint32_t* ownColor = tileColor ? tileColor + tid * S : p.color + ownPix * S;
  __syncthreads();

  for(uint32_t base = listBegin; base < listEnd; base += Q_CHUNK)
  {
    // ---- stage: one triangle per thread (the first Q_CHUNK threads) ---------------------------------------------------
    uint32_t nItems = 0;
    if(tid < Q_CHUNK && base + tid < listEnd)
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
      if(tid <= Q_CHUNK)
        itemStart[tid] = ex;  // thread Q_CHUNK holds the total (the threads behind the chunk contribute nothing)
    }
    __syncthreads();

    for(uint32_t k0 = 0; k0 < total; k0 += Q_BATCH)
    {
      const uint32_t par      = parity & 1u;
      uint32_t*      setWords = reinterpret_cast<uint32_t*>(pixSet + par * TILE_PIX);
      uint32_t*      list     = lists;            // compact, unordered
      uint32_t*      sorted   = lists + Q_BATCH;  // pixel-major

      // ---- A: coverage; compact list of the covered candidates ---------------------------------------------------------------
      if(tid == 0)
        sCount[par ^ 1u] = 0u;
      {
        uint32_t recs[Q_IPT];
        coverCandidates<S, Q_CHUNK>(p, slots, itemStart, k0 + tid * Q_IPT, total, tileX0, tileY0, yLocal0, setWords, recs);
        appendCovered(recs, &sCount[par], list);
      }
      __syncthreads();  // (1)

      // ---- B: thread = pixel: number of fragments of the pixel in this batch, their place in pixel-major order ----------------
      const uint32_t n = sCount[par];
      pixSet[(par ^ 1u) * TILE_PIX + tid] = make_uint4(0u, 0u, 0u, 0u);  // the next batch's sets (every reader is past barrier 1)
      const uint4    m  = pixSet[par * TILE_PIX + tid];
      const uint32_t c0 = __popc(m.x), c1 = __popc(m.y), c2 = __popc(m.z), c3 = __popc(m.w);
      const uint32_t c  = c0 + c1 + c2 + c3;
      uint32_t       off;
      {
        const uint32_t incl = warpInclusiveScan(c);
        if(lane == 31)
          warpTot[warp] = incl;
        __syncthreads();  // (2)
        uint32_t wb = 0;
#pragma unroll
        for(int v = 0; v < RASTER_THREADS / 32 - 1; v++)
          wb += v < warp ? warpTot[v] : 0u;
        off = wb + incl - c;
      }
      if(c)
      {
        pixOff[tid] = off;
        pixPre[tid] = (c0 << 8) | ((c0 + c1) << 16) | ((c0 + c1 + c2) << 24);
      }
      __syncthreads();  // (3)

      // ---- P: every fragment to its pixel-major position ------------------------------------------------------------------------
      for(uint32_t i = tid; i < n; i += RASTER_THREADS)
      {
        const uint32_t rec  = list[i];
        const uint32_t slot = rec & (Q_CHUNK - 1), pl = (rec >> 8) & 255u, w = slot >> 5;
        const uint32_t rank = __popc(setWords[pl * 4 + w] & ((1u << (slot & 31u)) - 1u)) + ((pixPre[pl] >> (8u * w)) & 255u);
        sorted[pixOff[pl] + rank] = rec;
      }
      __syncthreads();  // (4)

      for(uint32_t f0 = 0; f0 < n; f0 += FRAGS_PER_ROUND)
      {
        const uint32_t nFr = min((uint32_t)FRAGS_PER_ROUND, n - f0);  // fragments of this round

        // ---- C: shade one invocation per thread, densely, into the round's queue ---------------------------------------------------
        {
          const uint32_t fi = (uint32_t)tid / ENTRIES;
          const int      sI = PERSAMPLE ? tid % ENTRIES : 0;
          if(fi < nFr)
          {
            const uint32_t rec = sorted[f0 + fi];
            if(!PERSAMPLE || ((rec >> (16 + sI)) & 1u))
            {
              const TriSlot& s     = slots[rec & (Q_CHUNK - 1)];
              const bool     small = (s.box >> 20) & 1u;
              const int      lx = (rec >> 8) & 15, ly = (rec >> 12) & 15;
              // varyings and gl_FragCoord.z at the pixel centre, or at the sample under sample shading
              const int  px = ((tileX0 + lx) << 8) + (PERSAMPLE ? SamplePattern<S>::xr(sI) : 128);
              const int  py = ((tileY0 + ly) << 8) + (PERSAMPLE ? SamplePattern<S>::yr(sI) : 128);
              const Bary b  = makeBary(edgeFloat(s, 1, px, py, small), edgeFloat(s, 2, px, py, small), s.rarea);
              if(PASS == PASS_LOOP_DEPTH)
                qDepth[tid] = depthAt(s, b);
              else
              {
                float  vz   = 0.f;
                Color4 rgba = shadeAt<WEIGHTED>(p, s, b, vz);
                if(PASS == PASS_OPAQUE)
                {
                  // opaque.frag.glsl:30-34, BlendMode::NONE (main.cpp:541-546): alpha 1, stored as is
                  rgba.a      = 1.0f;
                  qColor[tid] = make_float4(__uint_as_float(encodeDst(tabs, rgba)), 0.f, 0.f, 0.f);
                }
                else
                {
                  qColor[tid] = make_float4(rgba.r, rgba.g, rgba.b, rgba.a);
                  qDepth[tid] = WEIGHTED ? vz : depthAt(s, b);
                }
              }
            }
          }
        }
        __syncthreads();

        // ---- D: inserts in primitive order ---------------------------------------------------------------------------------------
        if(PERSAMPLE)
        {
          // Sample shading: every (pixel, sample) has its own A-buffer list, aux words and colour sample, so every (pixel,
          // sample) is its own chain.  The round's fragments are pixel-major: a pixel's fragments are a RUN, and the thread
          // that finds the start of a run (for its sample) owns the chain for this round -- S times the owners of a per-pixel walk.
          for(uint32_t ch = tid; ch < nFr * S; ch += RASTER_THREADS)
          {
            const uint32_t fi = ch / S;
            const int      sI = (int)(ch % S);
            const uint32_t pl = (sorted[f0 + fi] >> 8) & 255u;
            if(fi > 0 && ((sorted[f0 + fi - 1] >> 8) & 255u) == pl)
              continue;  // not the first fragment of its pixel in this round
            const int    gx = tileX0 + (int)(pl & 15u), yl = yLocal0 + (int)(pl >> 4);
            uint32_t*    px = (tileColor ? tileColor + pl * S : p.color + ((size_t)yl * p.W + gx) * S) + sI;
            for(uint32_t g = fi; g < nFr; g++)
            {
              const uint32_t rec = sorted[f0 + g];
              if(((rec >> 8) & 255u) != pl)
                break;
              if(!((rec >> (16 + sI)) & 1u))
                continue;
              const float4 v   = qColor[g * S + sI];
              const Color4 out = invokeOwned<PASS, S>(ctx, gx, yl, (uint32_t)sI, 1u << sI, Color4{v.x, v.y, v.z, v.w}, qDepth[g * S + sI], 0.f, (int)pl);
              if(!isZero(out))
                *px = ropPremult(tabs, *px, out);
            }
          }
          __syncthreads();
        }
        else
        {
          // thread = pixel owner: walks ITS fragments of the round.  With coverage masks (S > 1) the ROP tail blend -- a decode,
          // blend and sRGB encode per covered sample -- is not done here: the owner leaves the colour in the queue and phase E
          // blends it per sample.
          constexpr bool DEFER_ROP = S > 1 && PASS != PASS_OPAQUE && PASS != PASS_WEIGHTED && PASS != PASS_LOOP_DEPTH;
          bool           anyRop    = false;
          if(c)
          {
            const uint32_t qa = max(off, f0), qb = min(off + c, f0 + nFr);
            for(uint32_t q = qa; q < qb; q++)
            {
              const uint32_t rec  = sorted[q];
              const uint32_t mask = (rec >> 16) & 255u;
              const uint32_t e    = q - f0;
              if(PASS == PASS_OPAQUE)
              {
                // depth test LESS + depth write per covered sample (main.cpp:530-532); the early test of phase A ran before
                // this pixel's earlier fragments of the batch were written: test again
                const TriSlot& s     = slots[rec & (Q_CHUNK - 1)];
                const bool     small = (s.box >> 20) & 1u;
                const uint32_t enc   = __float_as_uint(qColor[e].x);
                bool           wrote = false;
#pragma unroll 1
                for(int sd = 0; sd < S; sd++)
                  if(mask & (1u << sd))
                  {
                    const int   px = (ownX << 8) + SamplePattern<S>::xr(sd), py = ((tileY0 + ownLy) << 8) + SamplePattern<S>::yr(sd);
                    const Bary  b  = makeBary(edgeFloat(s, 1, px, py, small), edgeFloat(s, 2, px, py, small), s.rarea);
                    const float zs = depthAt(s, b);
                    if(zs < p.depth[ownPix * S + sd])
                    {
                      p.depth[ownPix * S + sd] = zs;
                      p.color[ownPix * S + sd] = enc;
                      wrote                    = true;
                    }
                  }
                ctx.nOpaque += wrote ? 1u : 0u;
              }
              else
              {
                const float4 v   = PASS == PASS_LOOP_DEPTH ? make_float4(0.f, 0.f, 0.f, 0.f) : qColor[e];
                const float  zq  = qDepth[e];
                const Color4 out = invokeOwned<PASS, S>(ctx, ownX, yLocal0 + ownLy, 0u, mask, Color4{v.x, v.y, v.z, v.w}, zq, zq, tid);
                if(DEFER_ROP)
                {
                  const bool rop = !isZero(out);
                  qDepth[e]      = rop ? 1.0f : 0.0f;
                  if(rop)
                    qColor[e] = make_float4(out.r, out.g, out.b, out.a);
                  anyRop = anyRop || rop;
                }
                else if(PASS != PASS_LOOP_DEPTH && PASS != PASS_WEIGHTED && !isZero(out))
                  ownColor[0] = ropPremult(tabs, ownColor[0], out);
              }
            }
          }
          if(DEFER_ROP)
          {
            // ---- E: the ROP of the round's tail blends, one chain per (pixel, SAMPLE) ------------------------------------------------
            if(__syncthreads_or(anyRop ? 1 : 0))
            {
              for(uint32_t ch = tid; ch < nFr * S; ch += RASTER_THREADS)
              {
                const uint32_t fi = ch / S;
                const int      sI = (int)(ch % S);
                const uint32_t pl = (sorted[f0 + fi] >> 8) & 255u;
                if(fi > 0 && ((sorted[f0 + fi - 1] >> 8) & 255u) == pl)
                  continue;  // not the first fragment of its pixel in this round
                uint32_t* px =
                    (tileColor ? tileColor + pl * S : p.color + ((size_t)(yLocal0 + (int)(pl >> 4)) * p.W + tileX0 + (int)(pl & 15u)) * S) + sI;
                for(uint32_t g = fi; g < nFr; g++)
                {
                  const uint32_t rec = sorted[f0 + g];
                  if(((rec >> 8) & 255u) != pl)
                    break;
                  if(((rec >> (16 + sI)) & 1u) && qDepth[g] != 0.0f)
                  {
                    const float4 v = qColor[g];
                    *px            = ropPremult(tabs, *px, Color4{v.x, v.y, v.z, v.w});
                  }
                }
              }
              __syncthreads();
            }
          }
          else
            __syncthreads();
        }
      }
      parity++;
    }
    __syncthreads();  // the slots are about to be replaced
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
      if(ownValid)
      {
        const AbufView av{ctx.abuf, ctx.aux, ctx.viewSize};
        fusedCompositePixel<S, passAlgorithm(PASS)>(p, tabs, A, tid, av, ctx.onChip ? (size_t)tid : ownPix, ownPix, tileColorSm + tid * S,
                                                    WEIGHTED ? wAccSm + tid * S : nullptr, WEIGHTED ? wRevSm + tid * S : nullptr);
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
    if(lane == 0 && v)
      atomicAdd(&p.stats[q == 0 ? STAT_FRAGMENTS : (q == 1 ? STAT_STORED : (q == 2 ? STAT_TAIL : STAT_OPAQUE))], (unsigned long long)v);
  }
}

// ---- dispatch ----------------------------------------------------------------------------------------------------------
template <int PASS, int S, bool SSHADE>
static void launchKernelQ(const FrameParams& p, unsigned grid, cudaStream_t s)
{
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
    if(cudaFuncSetAttribute(k_raster_q<PASS, S, SSHADE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(TILE_PIX * S * 10 + ON_CHIP_MAX_BYTES))
       != cudaSuccess)
      return;  // stays in cudaGetLastError(), which every stage entry point checks after its launches
    if(dev >= 0 && dev < 64)
      configured[dev] = true;
  }
  const bool fused = p.fused && PASS != PASS_OPAQUE && PASS != PASS_LOOP_DEPTH;
  k_raster_q<PASS, S, SSHADE><<<grid, RASTER_THREADS, fused ? dynBytes : 0, s>>>(p);
}

template <int PASS>
static void launchPassQ(const FrameParams& p, cudaStream_t s)
{
  const unsigned grid = (unsigned)(p.tilesX * p.tileRowsLocal);
  const bool     ss   = p.sampleShading != 0;
  if(p.msaa == 1)
    launchKernelQ<PASS, 1, false>(p, grid, s);
  else if(p.msaa == 4)
  {
    if(ss)
      launchKernelQ<PASS, 4, true>(p, grid, s);
    else
      launchKernelQ<PASS, 4, false>(p, grid, s);
  }
  else
  {
    if(ss)
      launchKernelQ<PASS, 8, true>(p, grid, s);
    else
      launchKernelQ<PASS, 8, false>(p, grid, s);
  }
}

int launchRasterQueued(const FrameParams& p, int pass, cudaStream_t s)
{
  if(p.tilesX * p.tileRowsLocal == 0)
    return 0;
  switch(pass)
  {
    case PASS_SIMPLE: launchPassQ<PASS_SIMPLE>(p, s); break;
    case PASS_LINKEDLIST: launchPassQ<PASS_LINKEDLIST>(p, s); break;
    case PASS_LOOP_COLOR: launchPassQ<PASS_LOOP_COLOR>(p, s); break;
    case PASS_LOOP64: launchPassQ<PASS_LOOP64>(p, s); break;
    case PASS_SPINLOCK: launchPassQ<PASS_SPINLOCK>(p, s); break;
    case PASS_INTERLOCK: launchPassQ<PASS_INTERLOCK>(p, s); break;
    case PASS_WEIGHTED: launchPassQ<PASS_WEIGHTED>(p, s); break;
    case PASS_LOOP_DEPTH: launchPassQ<PASS_LOOP_DEPTH>(p, s); break;
    case PASS_OPAQUE: launchPassQ<PASS_OPAQUE>(p, s); break;
    default: return 0;
  }
  return 1;
}

}  // namespace oit
