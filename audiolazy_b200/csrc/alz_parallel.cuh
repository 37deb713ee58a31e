// alz_parallel.cuh -- ParallelFilter in ONE kernel (reference lazy_filters.py:1048-1054: every member filters
// the same input, the outputs are summed left to right).
//
// CTA = one warp = 32 streams; the warp walks over ALL channels of the plan for each tile of 32 samples:
//   acc[j] = y_0[j];  acc[j] = acc[j] + y_c[j]  for c = 1 .. C-1          (float64 registers)
// and writes float32(acc) once.  The channel outputs never exist in memory: 4 B read + 4 B written per input
// sample instead of 4 + 8 C (round 1: bank launch + a second kernel that read the C float32 rows back), and the sum
// is taken over the float64 channel results, as the reference's, before the single rounding to float32.
// A channel's recurrence state lives in the state buffer between tiles (it cannot stay in registers while the other
// channels run): (K+1)*2+ doubles loaded and stored per channel and tile, coalesced and L2 resident.
#pragma once
#include "alz_lane_tma.cuh"

template <class Core, class CoreArgs>
__device__ __forceinline__ void alz_run_warp_parallel(const AlzTileArgs& a, const CoreArgs& ca, const CUtensorMap* tmx,
                                                      const CUtensorMap* tmo, unsigned char* smem) {
  const int lane = threadIdx.x;
  const long long s0 = (long long)blockIdx.x * 32;
  const long long s = s0 + lane;
  const bool valid = s < a.S;
  const long long srow = valid ? s : a.S - 1;
  const unsigned tile0 = alz_smem_u32(smem);
  const unsigned mbar0 = tile0 + 2 * ALZ_TMA_TILE_BYTES;
  if (lane == 0) {
    alz_mbar_init(mbar0, 1);
    alz_mbar_init(mbar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncwarp();
  const int ntiles = (int)((a.T + ALZ_TT - 1) / ALZ_TT);
  const int swz = lane & 7;
  float* const myrow = reinterpret_cast<float*>(smem) + lane * 32;
  const bool tail_by_lanes = (a.T & 3) != 0;
  if (lane == 0) {
    alz_mbar_expect_tx(mbar0, ALZ_TMA_TILE_BYTES);
    alz_tma_load_2d(tile0, tmx, 0, (int)s0, mbar0);
  }
  for (int i = 0; i < ntiles; ++i) {
    const int b = i & 1, t0 = i * ALZ_TT;
    if (lane == 0 && i + 1 < ntiles) {
      if (i >= 1) alz_bulk_wait_read0();               // the other buffer was the source of tile i-1's store
      alz_mbar_expect_tx(mbar0 + 8 * (b ^ 1), ALZ_TMA_TILE_BYTES);
      alz_tma_load_2d(tile0 + (b ^ 1) * ALZ_TMA_TILE_BYTES, tmx, t0 + ALZ_TT, (int)s0, mbar0 + 8 * (b ^ 1));
    }
    alz_mbar_wait(mbar0 + 8 * b, (i >> 1) & 1);
    const int nvalid = (int)(a.T - t0 < ALZ_TT ? a.T - t0 : ALZ_TT);
    float* row = myrow + b * (ALZ_TMA_TILE_BYTES / 4);
    double acc[ALZ_TT];
#pragma unroll 1
    for (int pos = 0; pos < ca.n_pos; ++pos) {
      Core core;
      const long long r = (long long)ca.channel(pos) * a.Stot + srow;
      core.load(a, ca, r, pos, valid);
      core.template acc_from<0>(row, swz, nvalid, pos == 0, acc);
      if (valid) core.store(a, r, nvalid);
    }
#pragma unroll
    for (int g = 0; g < ALZ_TT / 4; ++g) {
      float4 o;
      o.x = (float)acc[4 * g + 0]; o.y = (float)acc[4 * g + 1]; o.z = (float)acc[4 * g + 2]; o.w = (float)acc[4 * g + 3];
      *reinterpret_cast<float4*>(row + ((g ^ swz) << 2)) = o;
    }
    alz_fence_async_smem();
    __syncwarp();
    const bool by_lanes = tail_by_lanes && i + 1 == ntiles;
    if (by_lanes) {
      if (valid) {
        float* dst = a.y + s * a.ys + t0;
        for (int j = 0; j < nvalid; ++j) dst[j] = row[(((j >> 2) ^ swz) << 2) | (j & 3)];
      }
    } else if (lane == 0) {
      alz_tma_store_2d(tmo, t0, (int)s0, tile0 + b * ALZ_TMA_TILE_BYTES);
    }
    if (lane == 0) alz_bulk_commit();
  }
  if (lane == 0) alz_bulk_wait0();
}
