// Biquad kernels for cascades of K = 1 sections (see alz_launch.cuh).
#include "alz_launch.cuh"
int alzi_launch_biquad_k1(const alz_plan* p, const AlzTileArgs& ta, cudaStream_t st) { return launch_biquad_k<1>(p, ta, st); }
double alzi_probe_biquad_k1(const alz_plan* p, const double* r64, const double* r32) { return probe_biquad_k<1>(p, r64, r32); }
