// ParallelFilter kernels (alz_parallel.cuh): all channels of a plan summed inside one kernel.
#include "alz_biquad.cuh"
#include "alz_parallel.cuh"
#include "alz_plan.h"

static const int kCoefSmall = 512, kCoefLarge = 3584;   // as alz_launch.cuh

template <int K, int NCOEF>
__global__ void __launch_bounds__(32, 12)
alz_parallel_sum_kernel(const __grid_constant__ AlzTileArgs a, const __grid_constant__ AlzBiquadArgs<NCOEF> ca,
                        const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmo) {
  extern __shared__ __align__(1024) unsigned char alz_smem_tma[];
  alz_run_warp_parallel<AlzBiquadCore<K, 3, 0, 0, 0, double>>(a, ca, &tmx, &tmo, alz_smem_tma);
}

template <int K>
static int launch_parallel_k(const alz_plan* p, const AlzTileArgs& ta, const CUtensorMap& tmx, const CUtensorMap& tmo, cudaStream_t st) {
  const unsigned groups = (unsigned)((ta.S + 31) / 32);
  void* args[4] = {(void*)&ta, p->chunks[0].block, (void*)&tmx, (void*)&tmo};
  const void* kern = p->coef_small ? (const void*)alz_parallel_sum_kernel<K, kCoefSmall> : (const void*)alz_parallel_sum_kernel<K, kCoefLarge>;
  ALZ_CUDA(cudaLaunchKernel(kern, dim3(groups), dim3(32), args, ALZ_TMA_SMEM, st));
  ALZ_CUDA(cudaGetLastError());
  alzi_launches.fetch_add(1, std::memory_order_relaxed);
  return ALZI_OK;
}

int alzi_launch_parallel(const alz_plan* p, const AlzTileArgs& ta, const CUtensorMap& tmx, const CUtensorMap& tmo, cudaStream_t st) {
  switch (p->K) {
    case 1: return launch_parallel_k<1>(p, ta, tmx, tmo, st);
    case 2: return launch_parallel_k<2>(p, ta, tmx, tmo, st);
    case 3: return launch_parallel_k<3>(p, ta, tmx, tmo, st);
    case 4: return launch_parallel_k<4>(p, ta, tmx, tmo, st);
    case 6: return launch_parallel_k<6>(p, ta, tmx, tmo, st);
    case 8: return launch_parallel_k<8>(p, ta, tmx, tmo, st);
  }
  return alzi_fail(ALZI_ERR_UNSUPPORTED, "no parallel-sum kernel for K=%d", p->K);
}
