// Biquad kernels for cascades of K = 4 sections (see alz_launch.cuh).
#include "alz_launch.cuh"
int alzi_launch_biquad_k4(const alz_plan* p, const AlzTileArgs& ta, cudaStream_t st) { return launch_biquad_k<4>(p, ta, st); }
double alzi_probe_biquad_k4(const alz_plan* p, const double* r64, const double* r32) { return probe_biquad_k<4>(p, r64, r32); }
int alzi_launch_envelope_k4(const alz_plan* p, const AlzTileArgs& ta, cudaStream_t st) {
  if (p->NB <= 2) return launch_envelope_t<4, 2, 0, 0>(p, ta, st);
  if ((p->zmask & ALZ_ZMASK_KLAPURI) == ALZ_ZMASK_KLAPURI) return launch_envelope_t<4, 3, 0, ALZ_ZMASK_KLAPURI>(p, ta, st);
  return launch_envelope_t<4, 3, 0, 0>(p, ta, st);
}
