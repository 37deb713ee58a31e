// alz_generic.cuh -- arbitrary-order / sparse-tap direct-form-I core (fallback kernel).
//
// Covers every LTI filter the reference's evaluator accepts that is not a cascade of
// biquads: long FIR numerators (gammatone.sampled's 8-tap first section,
// lazy_auditory.py:151-182), high-order single sections (lpc, maverage), sparse
// combs with large delays (comb.fb / comb.ff, lazy_filters.py:1087-1173).
//
// Same expression order as the reference's generated source (lazy_filters.py:197-237):
// numerator taps by ascending delay, then denominator taps by ascending delay, with
// a0 folded into the coefficients on the host and the products fused (DFMA).
//
// Histories live in global memory (the state buffer), one power-of-two ring per
// section for the input and one for the output, indexed by the ABSOLUTE sample count
// so that block splitting is bit-exact.  Layout state[slot * sstride + r], r = c*Stot + s.
// The tap structure (delays) is the union over channels, the coefficients are read
// with warp-uniform addresses coef[tap * C + c] (all lanes of a warp share the channel).
#pragma once
#include "alz_lane.cuh"

struct AlzGenSection {
  int num_begin, nnum;   // taps [num_begin, num_begin+nnum): numerator, ascending delay
  int den_begin, nden;   // denominator taps (delay >= 1), coefficient already negated
  int xbase, xmask;      // input ring: slots [xbase, xbase+xmask+1), xmask = -1 if none
  int ybase, ymask;      // output ring
};

struct AlzGenericArgs {
  const AlzGenSection* sec;   // [K]
  const int* tap_delay;       // [ntaps]
  const double* coef;         // [ntaps][C]
  int K;
  int C;
  int c_base;
  // time-varying coefficients (reference lazy_filters.py:200-216): when non-null, tap i of
  // sample j of THIS launch uses tv[i * tv_stride + j] instead of coef[i][c] (C must be 1)
  const double* tv;
  long long tv_stride;
  __device__ __forceinline__ int channel(int pos) const { return c_base + pos; }
};

struct AlzGenericCore {
  const AlzGenSection* sec;
  const int* tap_delay;
  const double* cf;
  double* st;
  long long R;
  long long cnt0;
  int K, C;
  const double* tv;
  long long tv_stride;
  bool live;   // lanes beyond the last stream must not touch the (clamped) state rows

  __device__ __forceinline__ void load(const AlzTileArgs& a, const AlzGenericArgs& ca, long long r, int c_local, bool valid) {
    sec = ca.sec;
    tap_delay = ca.tap_delay;
    K = ca.K;
    C = ca.C;
    R = a.sstride;               // state slot 0 = absolute sample count
    cf = ca.coef + (ca.c_base + c_local);
    st = a.state + r;
    cnt0 = (long long)st[0];
    tv = ca.tv;
    tv_stride = ca.tv_stride;
    live = valid;
  }

  __device__ __forceinline__ double coef_at(int tap, long long j) const {
    return tv ? tv[(long long)tap * tv_stride + j] : cf[(long long)tap * C];
  }

  __device__ __forceinline__ float step(double xin, long long n, long long j) {
    double in = xin;
    for (int k = 0; k < K; ++k) {
      const AlzGenSection s = sec[k];
      double acc = 0.0;
      for (int i = 0; i < s.nnum; ++i) {
        const int d = tap_delay[s.num_begin + i];
        const double v = d == 0 ? in : st[(long long)(s.xbase + (int)((n - d) & s.xmask)) * R];
        acc = fma(coef_at(s.num_begin + i, j), v, acc);
      }
      for (int i = 0; i < s.nden; ++i) {
        const int d = tap_delay[s.den_begin + i];
        const double v = st[(long long)(s.ybase + (int)((n - d) & s.ymask)) * R];
        acc = fma(coef_at(s.den_begin + i, j), v, acc);
      }
      if (live) {
        if (s.xmask >= 0) st[(long long)(s.xbase + (int)(n & s.xmask)) * R] = in;
        if (s.ymask >= 0) st[(long long)(s.ybase + (int)(n & s.ymask)) * R] = acc;
      }
      in = acc;
    }
    return (float)in;
  }

  __device__ __forceinline__ void tile(float* row, int swz, int nvalid, long long n_done) {
    for (int j = 0; j < nvalid; ++j) {
      float* p = row + ((((j >> 2) ^ swz) << 2) | (j & 3));
      *p = step((double)*p, cnt0 + n_done + j, n_done + j);
    }
  }

  __device__ __forceinline__ void store(const AlzTileArgs&, long long, long long T) { st[0] = (double)(cnt0 + T); }
};
