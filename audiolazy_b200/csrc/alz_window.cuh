// alz_window.cuh -- single-section direct form I of ANY order and sparsity, register resident.
//
// Covers what is not a cascade of biquads but is one section per channel (reference
// LinearFilter.__call__, lazy_filters.py:197-257, with the filters that reach it from
// lazy_filters.py:1087-1173 comb.fb/.tau/.ff, lazy_synth.py:624-657 karplus_strong,
// lazy_lpc.py:142-340 analysis FIR / all-pole synthesis, lazy_analysis.py:569-616 maverage):
//     y[n] = sum_i b_i x[n-i] + sum_{i>=1} (-a_i) y[n-i]          (a0 folded in on the host)
// The taps are split by delay against the block length B = 16:
//   * NEAR taps, delay 1 .. M-1 (M = 4 or 16 slots): a DENSE window per history held in
//     registers.  The window is a ring indexed by the sample's position inside the block; blocks
//     are fully unrolled, so every index is a compile-time constant and nothing is ever shifted.
//     At a block start the value of delay i sits in slot (M - i) & (M - 1) ("canonical"); after a
//     whole block (B is a multiple of M) the ring is canonical again.
//   * FAR taps, delay >= 16 (any number, any delay: combs, karplus_strong, long FIRs): their
//     operands lie entirely BEFORE the block, in a power-of-two ring of the state buffer indexed
//     by the absolute sample count (slot-major: the 32 lanes of a warp read one 256-byte run).  All
//     16 samples of a block are independent of each other through these taps, so each tap is 16
//     independent coalesced loads + 16 DFMAs issued ahead of the serial part.
// Per sample the serial dependency is ONE DFMA (the delay-1 feedback tap is accumulated last).
// Coefficients of the near window come from the kernel-parameter constant bank (uniform operands),
// far-tap coefficients from a small global table read with warp-uniform addresses.
// Accumulation order differs from the reference's (far taps, numerator, denominator by DEscending
// delay): float64 reassociation, 1e-16 relative, 11 orders of magnitude inside the parity bar.
#pragma once
#include "alz_lane.cuh"

#define ALZ_WIN_BLOCK 16            // samples per block = far-tap threshold
#define ALZ_WIN_REC 32              // doubles per channel record: b[0..15], -a[1..15], pad

template <int NCOEF>
struct AlzWindowArgs {
  int n_far_x, n_far_y;             // far taps per history (delay >= 16)
  int xbase, xmask, ybase, ymask;   // far rings in the state buffer (mask = -1: none); slot 0 = absolute sample count
  int xwin, ywin;                   // first state slot of the near windows (delays 1 .. M-1)
  int C, pad_;
  const int* far_delay;             // [n_far_x + n_far_y]: numerator taps first
  const double* far_coef;           // [n_far_x + n_far_y][C] (denominator coefficients already negated)
  double coef[NCOEF];               // [C][ALZ_WIN_REC]
  __device__ __forceinline__ int channel(int pos) const { return pos; }
};

template <int MX, int MY, int NCOEF>   // window slots of the x / y history: 0 (no near taps), 4 or 16
struct AlzWindowCore {
  typedef AlzWindowArgs<NCOEF> Args;
  const Args* ca;
  const double* rec;                // this channel's coefficient record (constant bank)
  double* st;
  long long R, cnt0;
  int c;
  bool live;                        // lanes beyond the last stream must not touch the (clamped) state rows
  double xw[MX ? MX : 1], yw[MY ? MY : 1];

  __device__ __forceinline__ void load(const AlzTileArgs& a, const Args& args, long long r, int c_local, bool valid) {
    ca = &args;
    c = c_local;
    rec = args.coef + c_local * ALZ_WIN_REC;
    R = a.sstride;
    st = a.state + r;
    cnt0 = (long long)st[0];
    live = valid;
#pragma unroll
    for (int i = 1; i < MX; ++i) xw[(MX - i) & (MX - 1)] = st[(long long)(args.xwin + i - 1) * R];
#pragma unroll
    for (int i = 1; i < MY; ++i) yw[(MY - i) & (MY - 1)] = st[(long long)(args.ywin + i - 1) * R];
    if (MX) xw[0] = 0.0;
    if (MY) yw[0] = 0.0;
  }

  // One sample at block position j (compile-time): near taps from the rings, `far` = the far taps' sum.
  template <int J>
  __device__ __forceinline__ double sample(double x, double far) {
    double v = fma(rec[0], x, far);
#pragma unroll
    for (int i = MX - 1; i >= 1; --i) v = fma(rec[i], xw[(J - i) & (MX - 1)], v);
#pragma unroll
    for (int i = MY - 1; i >= 1; --i) v = fma(rec[16 + i - 1], yw[(J - i) & (MY - 1)], v);   // delay 1 last: one DFMA of serial latency
    if (MX) xw[J & (MX - 1)] = x;
    if (MY) yw[J & (MY - 1)] = v;
    return v;
  }

  template <int J>
  __device__ __forceinline__ void unrolled(const double (&xs)[ALZ_WIN_BLOCK], double (&acc)[ALZ_WIN_BLOCK]) {
    if constexpr (J < ALZ_WIN_BLOCK) {
      acc[J] = sample<J>(xs[J], acc[J]);
      unrolled<J + 1>(xs, acc);
    }
  }

  // Far taps of one history for `n` samples starting at absolute count n0: acc[j] += coef * ring[n0 + j - d].
  __device__ __forceinline__ void far_taps(double (&acc)[ALZ_WIN_BLOCK], long long n0, int first, int count, int base, int mask) {
    for (int f = first; f < first + count; ++f) {
      const int d = ca->far_delay[f];
      const double cf = ca->far_coef[(long long)f * ca->C + c];
      double v[ALZ_WIN_BLOCK];
#pragma unroll
      for (int j = 0; j < ALZ_WIN_BLOCK; ++j) v[j] = st[(long long)(base + (int)((n0 + j - d) & mask)) * R];
#pragma unroll
      for (int j = 0; j < ALZ_WIN_BLOCK; ++j) acc[j] = fma(cf, v[j], acc[j]);
    }
  }

  // A whole block of 16 samples; p = &row[first sample] (swizzled 16-byte chunks), n0 = its absolute count.
  __device__ __forceinline__ void block(float* row, int swz, int g0, long long n0) {
    double xs[ALZ_WIN_BLOCK], acc[ALZ_WIN_BLOCK];
#pragma unroll
    for (int g = 0; g < ALZ_WIN_BLOCK / 4; ++g) {
      const float4 xv = *reinterpret_cast<const float4*>(row + (((g0 + g) ^ swz) << 2));
      xs[4 * g + 0] = (double)xv.x; xs[4 * g + 1] = (double)xv.y; xs[4 * g + 2] = (double)xv.z; xs[4 * g + 3] = (double)xv.w;
    }
#pragma unroll
    for (int j = 0; j < ALZ_WIN_BLOCK; ++j) acc[j] = 0.0;
    if (ca->n_far_x) far_taps(acc, n0, 0, ca->n_far_x, ca->xbase, ca->xmask);
    if (ca->n_far_y) far_taps(acc, n0, ca->n_far_x, ca->n_far_y, ca->ybase, ca->ymask);
    unrolled<0>(xs, acc);
    if (live) {
      if (ca->xmask >= 0) {
#pragma unroll
        for (int j = 0; j < ALZ_WIN_BLOCK; ++j) st[(long long)(ca->xbase + (int)((n0 + j) & ca->xmask)) * R] = xs[j];
      }
      if (ca->ymask >= 0) {
#pragma unroll
        for (int j = 0; j < ALZ_WIN_BLOCK; ++j) st[(long long)(ca->ybase + (int)((n0 + j) & ca->ymask)) * R] = acc[j];
      }
    }
#pragma unroll
    for (int g = 0; g < ALZ_WIN_BLOCK / 4; ++g) {
      float4 o;
      o.x = (float)acc[4 * g + 0]; o.y = (float)acc[4 * g + 1]; o.z = (float)acc[4 * g + 2]; o.w = (float)acc[4 * g + 3];
      *reinterpret_cast<float4*>(row + (((g0 + g) ^ swz) << 2)) = o;
    }
  }

  // Ragged end of a launch (< 16 samples): one sample at a time with the rings kept canonical by shifting.
  __device__ __forceinline__ float single(float xin, long long n) {
    const double x = (double)xin;
    double far = 0.0;
    for (int f = 0; f < ca->n_far_x + ca->n_far_y; ++f) {
      const bool isy = f >= ca->n_far_x;
      const int base = isy ? ca->ybase : ca->xbase, mask = isy ? ca->ymask : ca->xmask;
      far = fma(ca->far_coef[(long long)f * ca->C + c], st[(long long)(base + (int)((n - ca->far_delay[f]) & mask)) * R], far);
    }
    const double v = sample<0>(x, far);          // position 0: reads slots (M - i), writes slot 0
    if (MX) {
#pragma unroll
      for (int k = 1; k < MX - 1; ++k) xw[k] = xw[k + 1];
      xw[MX - 1] = xw[0];
    }
    if (MY) {
#pragma unroll
      for (int k = 1; k < MY - 1; ++k) yw[k] = yw[k + 1];
      yw[MY - 1] = yw[0];
    }
    if (live) {
      if (ca->xmask >= 0) st[(long long)(ca->xbase + (int)(n & ca->xmask)) * R] = x;
      if (ca->ymask >= 0) st[(long long)(ca->ybase + (int)(n & ca->ymask)) * R] = v;
    }
    return (float)v;
  }

  __device__ __forceinline__ void tile(float* row, int swz, int nvalid, long long n_done) {
    const long long n0 = cnt0 + n_done;
    int j = 0;
    for (; j + ALZ_WIN_BLOCK <= nvalid; j += ALZ_WIN_BLOCK) block(row, swz, j >> 2, n0 + j);
    for (; j < nvalid; ++j) {
      float* p = row + ((((j >> 2) ^ swz) << 2) | (j & 3));
      *p = single(*p, n0 + j);
    }
  }

  __device__ __forceinline__ void store(const AlzTileArgs&, long long, long long T) {
    st[0] = (double)(cnt0 + T);
#pragma unroll
    for (int i = 1; i < MX; ++i) st[(long long)(ca->xwin + i - 1) * R] = xw[(MX - i) & (MX - 1)];
#pragma unroll
    for (int i = 1; i < MY; ++i) st[(long long)(ca->ywin + i - 1) * R] = yw[(MY - i) & (MY - 1)];
  }
};
