"""The plan-time precision-tier decision (host logic, no GPU): a design-only plan runs every channel
through the kernel's own arithmetic in float64 and float32 on the CPU.  Here the float32 error it reports
is re-derived with an independent numpy emulation of the same recurrence on the same probe noise."""
import numpy as np
import pytest

from audiolazy_b200 import _capi

f32 = np.float32


def lcg_noise(n, seed=12345):
  out = np.empty(n, dtype=f32)
  s = seed
  for i in range(n):
    s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
    out[i] = f32((s >> 8) * (2.0 / 16777216.0) - 1.0)
  return out


def fma32(a, b, c):
  return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def monic_emulation(bank, x, dtype, dform=None):
  """Monic cascade of csrc/alz_biquad.cuh (gain on the float32 input), vectorised over channels. ``dform``: boolean
  mask of the channels evaluated in the difference form (state y1, d1 = y1 - y2; coefficients h, e2)."""
  C, K = len(bank), max(len(ch) for ch in bank)
  c1 = np.zeros((C, K)); na1 = np.zeros((C, K)); na2 = np.zeros((C, K)); G = np.ones(C)
  for c, ch in enumerate(bank):
    for k, (b, a) in enumerate(ch):
      assert len(b) <= 2 and a[0] == 1.0
      G[c] *= b[0]
      c1[c, k] = (b[1] if len(b) > 1 else 0.0) / b[0]
      na1[c, k], na2[c, k] = -a[1], -a[2]
  Gf = G.astype(f32)
  fma = fma32 if dtype is f32 else (lambda a, b, c: a * b + c)     # float64: fused vs. separate rounding is far below what is compared
  dmask = np.zeros(C, dtype=bool) if dform is None else np.asarray(dform, dtype=bool)
  nh, ne2 = -(1.0 - na1 - na2), -(1.0 + na2)
  u = np.zeros((K + 1, 2, C), dtype=dtype)     # u[k+1][1] = y2 (direct form) or d1 (difference form)
  y = np.empty((C, len(x)), dtype=f32)
  cf = lambda v: v.astype(dtype)
  for n in range(len(x)):
    inp = (x[n] * Gf).astype(f32).astype(dtype)
    in1 = u[0, 0].copy()
    u[0, 1] = in1
    u[0, 0] = inp
    for k in range(K):
      y1, s2 = u[k + 1, 0].copy(), u[k + 1, 1].copy()
      t = fma(cf(c1[:, k]), in1, inp)
      o_dir = fma(cf(na1[:, k]), y1, fma(cf(na2[:, k]), s2, t))
      w = fma(cf(ne2[:, k]), s2, fma(cf(nh[:, k]), y1, t))
      d = (s2 + w).astype(dtype)
      o_dif = (y1 + d).astype(dtype)
      u[k + 1, 1] = np.where(dmask, d, y1)
      u[k + 1, 0] = np.where(dmask, o_dif, o_dir)
      inp, in1 = u[k + 1, 0].copy(), y1
    y[:, n] = inp.astype(f32)
  return y


def test_tier_probe_matches_an_independent_emulation(designs):
  bank = designs["bank_slaney"]
  plan = _capi.Plan(bank, design_only=True)
  assert plan.kind == _capi.KIND_BIQUAD and plan.device == -1
  tier, probe = plan.tiers()
  assert plan.n_fp32_channels == int((tier > 0).sum()) > 0 and plan.n_dform_channels == int((tier == 2).sum()) > 0
  assert np.all(probe[tier > 0] <= plan.tier_tol) and np.all(probe[tier == 0] > plan.tier_tol)
  x = lcg_noise(8192)
  y64 = monic_emulation(bank, x, np.float64).astype(np.float64)
  peak = np.max(np.abs(y64), axis=1)
  err_dir = np.max(np.abs(monic_emulation(bank, x, f32).astype(np.float64) - y64), axis=1) / peak
  err_dif = np.max(np.abs(monic_emulation(bank, x, f32, dform=np.ones(len(bank), bool)).astype(np.float64) - y64), axis=1) / peak
  # the probe also runs a step, an impulse and noise + Nyquist tone, so it may only be larger than the noise-only figure
  mine = np.where(tier == 2, err_dif, np.where(tier == 1, err_dir, np.minimum(err_dir, err_dif)))
  assert np.all(probe >= mine * 0.98)
  assert np.median(probe[tier > 0] / mine[tier > 0]) < 1.3
  # the picture: the direct form in float32 only holds for the upper channels, the difference form reaches far down,
  # the lowest channels stay in float64
  assert not tier[:4].any() and np.all(tier[-16:] == 1) and (tier == 2).sum() >= 20
  assert err_dir[:16].min() > 2e-5 and err_dif[12:40].max() < 2.5e-6
  strict = _capi.Plan(bank, design_only=True, strict_tiers=True)   # + the pure Nyquist sequence
  assert strict.n_fp32_channels < plan.n_fp32_channels and strict.n_dform_channels <= 4


def test_design_only_plans_cannot_compute(designs):
  plan = _capi.Plan(designs["bank_klapuri"], design_only=True)
  with pytest.raises(_capi.NativeError):
    plan.apply(0, 0, 0, 1, 1, 1, 1)
  exact = _capi.Plan(designs["bank_klapuri"], design_only=True, exact=True)
  assert exact.n_fp32_channels == 0 and plan.n_fp32_channels > 0
  assert plan.state_doubles(3) == exact.state_doubles(3)


def test_generic_plans_have_no_tiers(designs):
  plan = _capi.Plan([designs["comb_fb_37_0.8"]], design_only=True)
  assert plan.kind == _capi.KIND_GENERIC and plan.n_fp32_channels == 0
  tier, probe = plan.tiers()
  assert not tier.any() and np.all(probe < 0)
