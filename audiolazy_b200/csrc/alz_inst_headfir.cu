// Head-FIR biquad kernels: first section with up to 8 numerator taps (gammatone.sampled,
// reference lazy_auditory.py:151-182), K in {1, 4} (see alz_launch.cuh).
#include "alz_launch.cuh"
int alzi_launch_headfir_k1(const alz_plan* p, const AlzTileArgs& ta, cudaStream_t st) { return launch_headfir_k<1>(p, ta, st); }
int alzi_launch_headfir_k4(const alz_plan* p, const AlzTileArgs& ta, cudaStream_t st) { return launch_headfir_k<4>(p, ta, st); }
double alzi_probe_headfir_k1(const alz_plan* p, const double* r64, const double* r32) { return probe_headfir_k<1>(p, r64, r32); }
double alzi_probe_headfir_k4(const alz_plan* p, const double* r64, const double* r32) { return probe_headfir_k<4>(p, r64, r32); }
int alzi_launch_envelope_headfir_k4(const alz_plan* p, const AlzTileArgs& ta, cudaStream_t st) {
  if (p->NB <= 1) return launch_envelope_t<4, 1, 8, 0>(p, ta, st);
  return launch_envelope_t<4, 3, 8, 0>(p, ta, st);
}
