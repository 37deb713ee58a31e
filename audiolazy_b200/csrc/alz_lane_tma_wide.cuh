// alz_lane_tma_wide.cuh -- EXPERIMENTAL variant of the TMA tile engine with W warps per CTA.
//
// Same algorithm as alz_run_warp_tma (alz_lane_tma.cuh): every warp is an independent worker with
// its own two tiles and mbarriers, there is no CTA barrier.  W > 1 only amortises the 1 KB of
// shared memory the system reserves per CTA (W = 3: 9 CTAs x 3 warps = 27 warps per SM instead of
// 24).  To fit the 72-register budget of 27 warps per SM the long-lived values are narrower here
// (32-bit segment bounds, state index and flag address recomputed at the end): 66 registers, no
// spills.  Selected with ALZ_WARPS_PER_CTA=3; NOT YET RUN ON HARDWARE.  Kept in a separate
// function so that the measured single-warp engine keeps its exact code.
#pragma once
#include "alz_lane_tma.cuh"

// All warps of a CTA share the channel (blockIdx.x); warp w of CTA y takes stream group / ticket y*W + w.
template <class Core, class CoreArgs, int W>
__device__ __forceinline__ void alz_run_warps_tma_wide(const AlzTileArgs& a, const CoreArgs& ca, const CUtensorMap* tmx,
                                                 const CUtensorMap* tmy, unsigned char* smem) {
  const int lane = W == 1 ? threadIdx.x : (threadIdx.x & 31);
  const int warp = W == 1 ? 0 : (threadIdx.x >> 5);
  const int c_local = blockIdx.x;              // CTA-uniform: coefficients go to uniform registers
  const int c = a.c_base + c_local;
  int group = W == 1 ? blockIdx.y : blockIdx.y * W + warp, seg = 0;
  int tbeg = 0, tlen = (int)a.T;
  unsigned* flag = nullptr;
  if constexpr (W > 1) {
    if (a.nseg <= 1 && group >= a.groups) return;   // surplus warp of the last CTA
  }
  if (a.nseg > 1) {
    // Ticket order = start order within the channel, so the CTA that owns the previous segment
    // of my (channel, group) is already running or done: the wait below cannot deadlock.
    unsigned ticket = 0;
    if (lane == 0) ticket = atomicAdd(a.sync + c_local, 1u);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if constexpr (W > 1) {
      if (ticket >= (unsigned)a.groups * (unsigned)a.nseg) return;
    }
    seg = (int)(ticket / (unsigned)a.groups);
    group = (int)(ticket - (unsigned)seg * (unsigned)a.groups);
    tbeg = seg * (int)a.seg_len;
    tlen = (int)a.T - tbeg < (int)a.seg_len ? (int)a.T - tbeg : (int)a.seg_len;
    flag = a.sync + gridDim.x + (size_t)c_local * a.groups + group;
    if (seg > 0) {
      unsigned done;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(done) : "l"(flag) : "memory");
      } while (done < (unsigned)seg);
    }
  }
  const int s0 = group * 32;
  const long long s = (long long)s0 + lane;
  const bool valid = s < a.S;
  const long long r = (long long)c * a.Stot + (valid ? s : a.S - 1);   // stream-fastest: coalesced state access

  if constexpr (W > 1) smem += warp * (2 * ALZ_TMA_TILE_BYTES);   // tiles of all warps first (1024-byte aligned) ...
  const unsigned tile0 = alz_smem_u32(smem);
  const unsigned mbar0 = W == 1 ? tile0 + 2 * ALZ_TMA_TILE_BYTES                                   // ... mbarriers after them
                                : tile0 + (W - warp) * (2 * ALZ_TMA_TILE_BYTES) + warp * 16;
  if (lane == 0) {
    alz_mbar_init(mbar0, 1);
    alz_mbar_init(mbar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();

  Core core;
  core.load(ca, r, c_local, valid);

  const int ntiles = (tlen + ALZ_TT - 1) / ALZ_TT;
  const int nfull = tlen / ALZ_TT;
  const int tb = tbeg;
  const int swz = lane & 7;
  float* const myrow = reinterpret_cast<float*>(smem) + lane * 32;

  const bool paired = a.paired != 0;
  const bool tail_by_lanes = (a.T & 3) != 0 && nfull < ntiles;
  if (lane == 0 && !paired) {   // tile 0 in flight
    alz_mbar_expect_tx(mbar0, ALZ_TMA_TILE_BYTES);
    alz_tma_load_2d(tile0, tmx, tb, (int)s0, mbar0);
  }

  for (int i = 0; i < ntiles; ++i) {
    const int b = i & 1;
    const int t0 = i * ALZ_TT;
    if (lane == 0) {
      if (paired) {
        // Tiles go in pairs: both loads are issued together once the previous pair's stores have
        // been read out of shared memory, and both stores are issued back to back after the second
        // tile, so each output row receives 256 contiguous bytes at (nearly) the same time.
        if (b == 0) {
          if (i >= 2) alz_bulk_wait_read0();
          alz_mbar_expect_tx(mbar0, ALZ_TMA_TILE_BYTES);
          alz_tma_load_2d(tile0, tmx, tb + t0, (int)s0, mbar0);
          if (i + 1 < ntiles) {
            alz_mbar_expect_tx(mbar0 + 8, ALZ_TMA_TILE_BYTES);
            alz_tma_load_2d(tile0 + ALZ_TMA_TILE_BYTES, tmx, tb + t0 + ALZ_TT, (int)s0, mbar0 + 8);
          }
        }
      } else if (i + 1 < ntiles) {
        // Prefetch tile i+1 into the other buffer.  That buffer was the source of the TMA
        // store of tile i-1: wait until the store has finished READING it (it was issued a
        // whole barrier-wait ago, so this normally does not block).
        if (i >= 1) alz_bulk_wait_read0();
        alz_mbar_expect_tx(mbar0 + 8 * (b ^ 1), ALZ_TMA_TILE_BYTES);
        alz_tma_load_2d(tile0 + (b ^ 1) * ALZ_TMA_TILE_BYTES, tmx, tb + t0 + ALZ_TT, (int)s0, mbar0 + 8 * (b ^ 1));
      }
    }
    alz_mbar_wait(mbar0 + 8 * b, (i >> 1) & 1);     // tile i has landed (async proxy writes visible after the wait)
    const int nvalid = i < nfull ? ALZ_TT : (int)(tlen - t0);
    core.tile(myrow + b * (ALZ_TMA_TILE_BYTES / 4), swz, nvalid, t0);
    alz_fence_async_smem();                          // my generic-proxy writes -> visible to the TMA store
    __syncwarp();
    const bool by_lanes = tail_by_lanes && i + 1 == ntiles;   // ragged last tile: stored after the loop
    if (lane == 0) {
      if (!paired) {
        if (!by_lanes) alz_tma_store_3d(tmy, tb + t0, c, (int)s0, tile0 + b * ALZ_TMA_TILE_BYTES);
        alz_bulk_commit();
      } else if (b == 1 || i + 1 == ntiles) {
        if (b == 1) alz_tma_store_3d(tmy, tb + t0 - ALZ_TT, c, (int)s0, tile0);
        if (!by_lanes) alz_tma_store_3d(tmy, tb + t0, c, (int)s0, tile0 + b * ALZ_TMA_TILE_BYTES);
        alz_bulk_commit();
      }
    }
  }
  if (tail_by_lanes && valid) {
    // The TMA clips a box at 16-byte granularity: when n_samples is not a multiple of 4 the ragged
    // last tile is written by the lanes themselves (plain stores of the valid samples only).
    const int i = ntiles - 1, t0 = i * ALZ_TT, nvalid = (int)(tlen - t0);
    const float* src = myrow + (i & 1) * (ALZ_TMA_TILE_BYTES / 4);
    float* dst = a.y + s * a.ysS + (long long)c * a.ys + tbeg + t0;
    for (int j = 0; j < nvalid; ++j) dst[j] = src[(((j >> 2) ^ swz) << 2) | (j & 3)];
  }
  if (lane == 0) alz_bulk_wait0();                   // all output tiles are globally written before exit
  if (valid) core.store(ca, (long long)c * a.Stot + ((long long)group * 32 + lane), tlen);
  if (a.nseg > 1 && seg + 1 < a.nseg) {         // hand the state to the next segment
    flag = a.sync + gridDim.x + (size_t)c_local * a.groups + group;
    __threadfence();
    __syncwarp();
    if (lane == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(flag), "r"((unsigned)(seg + 1)) : "memory");
  }
}
