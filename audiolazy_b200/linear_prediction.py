"""Linear predictive coding: block statistics -> analysis filter ``A(z)`` (a FIR ZFilter).

A caller of the filter hot path: ``lpc(blk, order)`` returns the
whitening FIR filter, ``1 / lpc(blk, order)`` the all-pole synthesis filter, and applying either
to a signal goes through the CUDA kernels like any other ZFilter (orders above a biquad use the
generic ring kernel). Mirrors reference ``audiolazy/lazy_lpc.py`` (strategy names, ``error``
attribute, exceptions) and ``acorr`` / ``lag_matrix`` of ``lazy_analysis.py:277-342``; the
recursions here work on coefficient lists, not on filter algebra, so results agree with the
reference to rounding (tests: 1e-9 relative), not bit for bit.
"""
from __future__ import annotations

import cmath
import itertools as it

from .core import StrategyDict
from .filters import ZFilter

__all__ = ["ParCorError", "acorr", "lag_matrix", "toeplitz", "levinson_durbin", "lpc", "parcor",
           "parcor_stable", "lsf", "lsf_stable"]


class ParCorError(ZeroDivisionError):
  """A reflection (partial correlation) coefficient cannot be found (``lazy_lpc.py:37-41``)."""


def acorr(blk, max_lag=None):
  """Autocorrelation ``[sum_n blk[n] * blk[n + lag] for lag in 0..max_lag]`` of a block with
  a length; lags past the block give zeros (reference ``lazy_analysis.py:277-312``)."""
  blk = list(blk)
  if max_lag is None:
    max_lag = len(blk) - 1
  return [sum(blk[n] * blk[n + lag] for n in range(len(blk) - lag)) for lag in range(max_lag + 1)]


def lag_matrix(blk, max_lag=None):
  """Covariance-method lag matrix: cell ``[i][j] = sum_n blk[n - i] * blk[n - j]`` over the
  ``n`` that need no padding (reference ``lazy_analysis.py:315-342``)."""
  blk = list(blk)
  if max_lag is None:
    max_lag = len(blk) - 1
  elif max_lag >= len(blk):
    raise ValueError("Block length should be higher than order")
  span = range(max_lag, len(blk))
  return [[sum(blk[n - i] * blk[n - j] for n in span) for i in range(max_lag + 1)] for j in range(max_lag + 1)]


def toeplitz(vect):
  """Symmetric Toeplitz matrix (list of lists) from its first row (``lazy_lpc.py:44-49``)."""
  vect = list(vect)
  return [[vect[abs(i - j)] for i in range(len(vect))] for j in range(len(vect))]


def _quadratic_form(matrix, coefs):
  return sum(matrix(i, j) * ai * aj for i, ai in enumerate(coefs) for j, aj in enumerate(coefs))


def _fir(coefs, error):
  filt = ZFilter(list(coefs))
  filt.error = error
  return filt


def levinson_durbin(acdata, order=None):
  """Solve the Yule-Walker equations ``R a = -r`` for the predictor of the given order from
  autocorrelation lags; returns the analysis filter ``1 + a1 z^-1 + ...`` with the squared
  prediction error in ``.error`` (reference ``lazy_lpc.py:52-136``). O(order**2)."""
  acdata = list(acdata)
  if order is None:
    order = len(acdata) - 1
  elif order >= len(acdata):
    acdata = acdata + [0] * (order + 1 - len(acdata))
  a = [1]
  for m in range(1, order + 1):
    # prediction error of the order m-1 predictor, as the quadratic form a' R a
    err = _quadratic_form(lambda i, j: acdata[abs(i - j)], a)
    acc = sum(ai * acdata[m - i] for i, ai in enumerate(a))
    try:
      k = -acc / err
    except ZeroDivisionError:
      raise ParCorError("Can't find next PARCOR coefficient")
    padded = a + [0]
    a = [x + k * y for x, y in zip(padded, reversed(padded))]
  return _fir(a, _quadratic_form(lambda i, j: acdata[abs(i - j)], a))


lpc = StrategyDict("lpc")


@lpc.strategy("autocor", "acorr", "autocorrelation", "auto_correlation")
def lpc(blk, order=None):
  """Autocorrelation-method LPC: the least-squares solver for small orders, Levinson-Durbin
  above 100 (falling back when a reflection coefficient is undefined), ``lazy_lpc.py:142-183``."""
  blk = list(blk)
  if order is None:
    order = len(blk) - 1
  if order < 100:
    return lpc.nautocor(blk, order)
  try:
    return lpc.kautocor(blk, order)
  except ParCorError:
    return lpc.nautocor(blk, order)


def _least_squares(matrix, rhs):
  import numpy as np
  return (np.linalg.pinv(np.asarray(matrix, dtype=np.float64)) @ -np.asarray(rhs, dtype=np.float64)).tolist()


@lpc.strategy("nautocor", "nacorr", "nautocorrelation", "nauto_correlation")
def lpc(blk, order=None):
  """Autocorrelation method, normal equations solved with the pseudo-inverse (``lazy_lpc.py:186-225``)."""
  blk = list(blk)
  if order is None:
    order = len(blk) - 1
  acdata = acorr(blk, order)
  coeffs = _least_squares(toeplitz(acdata[:-1]), acdata[1:])
  return _fir([1] + coeffs, acdata[0] + sum(r * c for r, c in zip(acdata[1:], coeffs)))


@lpc.strategy("kautocor", "kacorr", "kautocorrelation", "kauto_correlation")
def lpc(blk, order=None):
  """Autocorrelation method through Levinson-Durbin (``lazy_lpc.py:228-272``).

  >>> filt = lpc.kautocor([-1, 0, 1, 0] * 4, 2)
  >>> filt.numerator, filt.error
  ([1.0, 0.0, 0.875], 1.875)
  """
  blk = list(blk)
  if order is None:
    order = len(blk) - 1
  return levinson_durbin(acorr(blk, order), order)


@lpc.strategy("covar", "cov", "covariance", "ncovar", "ncov", "ncovariance")
def lpc(blk, order=None):
  """Covariance method (no windowing assumption), pseudo-inverse solver (``lazy_lpc.py:275-294``)."""
  phi = lag_matrix(blk, order)
  coeffs = _least_squares([row[1:] for row in phi[1:]], [row[0] for row in phi[1:]])
  return _fir([1] + coeffs, phi[0][0] + sum(r * c for r, c in zip(phi[0][1:], coeffs)))


@lpc.strategy("kcovar", "kcov", "kcovariance")
def lpc(blk, order=None):
  """Covariance method solved greedily, one lattice-like stage per order, without NumPy: stage
  ``m`` adds ``k_m`` times the part of ``z^-m`` that is orthogonal (under the lag matrix) to
  the earlier stages' directions. Raises ``ValueError("Unstable filter")`` when a ``|k| >= 1``
  (reference ``lazy_lpc.py:297-340``)."""
  phi = lag_matrix(blk, order)
  order = len(phi) - 1
  size = order + 1

  def inner(a, b):
    return sum(phi[i][j] * ai * bj for i, ai in enumerate(a) if ai for j, bj in enumerate(b) if bj)

  def delay(m):
    return [1 if i == m else 0 for i in range(size)]

  a = delay(0)
  basis = [delay(1)]
  beta = [inner(basis[0], basis[0])]
  m = 1
  while True:
    try:
      k = -inner(a, delay(m)) / beta[m - 1]
    except ZeroDivisionError:
      raise ZeroDivisionError("Can't find next coefficient")
    if k >= 1 or k <= -1:
      raise ValueError("Unstable filter")
    a = [x + k * y for x, y in zip(a, basis[m - 1])]
    if m >= order:
      return _fir(a, inner(a, a))
    nxt = delay(m + 1)
    gamma = [inner(nxt, basis[q]) / beta[q] for q in range(m)]
    for q in range(m):
      nxt = [x - gamma[q] * y for x, y in zip(nxt, basis[q])]
    basis.append(nxt)
    beta.append(inner(nxt, nxt))
    m += 1


def _monic_fir(fir_filt):
  den = fir_filt.denominator
  if len(den) != 1:
    raise ValueError("Filter has feedback")
  num = list(fir_filt.numerator)
  if den[0] != 1:
    num = [c / den[0] for c in num]
  return num


def parcor(fir_filt):
  """Generator of the reflection coefficients of a FIR filter, highest order first (step-down
  recursion; reference ``lazy_lpc.py:343-395``)."""
  a = _monic_fir(fir_filt)
  for m in range(len(a) - 1, 0, -1):
    k = a[m]
    yield k
    try:
      a = [(x - k * y) / (1 - k ** 2) for x, y in zip(a, reversed(a))][:m]
    except ZeroDivisionError:
      raise ParCorError("Can't find next PARCOR coefficient")
    a[0] = 1


def parcor_stable(filt):
  """True when every reflection coefficient of the denominator is inside the unit circle
  (``lazy_lpc.py:398-425``)."""
  try:
    return all(abs(k) < 1 for k in parcor(ZFilter(filt.denpoly)))
  except ParCorError:
    return False


def lsf(fir_filt):
  """Line spectral frequencies (rad/sample) of a FIR filter: the angles of the roots of the
  palindromic / antipalindromic pair, interleaved starting from the lowest (``lazy_lpc.py:428-457``)."""
  import numpy as np
  a = _monic_fir(fir_filt) + [0]
  rev = a[::-1]
  angles = []
  for sign in (1, -1):
    poly = [x + sign * y for x, y in zip(a, rev)]
    angles.append(sorted(cmath.phase(r) for r in np.roots(poly[::-1])))
  return tuple(it.chain.from_iterable(zip(*sorted(angles))))


def lsf_stable(filt):
  """True when the LSFs of the two polynomials strictly alternate (``lazy_lpc.py:460-487``)."""
  data = lsf(ZFilter(filt.denpoly))
  return all(x < y for x, y in zip(data, data[1:]))
