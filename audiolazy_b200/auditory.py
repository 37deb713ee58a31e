"""Auditory filterbank design: ERB bandwidths and gammatone cascades.

Host-side (float64) restatement of the filter-design half of reference
``audiolazy/lazy_auditory.py``: ``erb`` (``:34-88``), ``gammatone_erb_constants``
(``:91-125``) and the three ``gammatone`` strategies (``:128-218``), each returning a
4-section :class:`~audiolazy_b200.filters.CascadeFilter`, with the reference's
formulas in the same floating-point operation order (designs agree to the last bit
on the golden grid of ``tests/golden``). ``phon2dB`` (ISO-226 loudness tables) is not
filtering and is out of scope.

The reference has no filterbank object (its only bank usage is the loop of
``examples/gammatone_plots.py:42-73``); :func:`erb_space` and :func:`gammatone_bank`
define the 64-channel ERB bank of the benchmark configs on top of
:class:`~audiolazy_b200.bank.FilterBank`.
"""
from __future__ import annotations

from math import cos, exp, factorial, log10, pi, sin, sqrt

from .core import StrategyDict
from .filters import CascadeFilter, ZFilter, resonator, z
from .misc import elementwise, sHz

__all__ = ["erb", "gammatone", "gammatone_erb_constants", "erb_space", "gammatone_bank"]

erb = StrategyDict("erb")


def _in_hertz(freq, Hz):
  """``(frequency in Hz, unit)``; without a unit the argument must already be in Hz."""
  if Hz is None:
    if freq < 7:   # looks like rad/sample (anything up to 2 * pi)
      raise ValueError("Frequency out of range.")
    return freq / 1, 1
  return freq / Hz, Hz


@erb.strategy("gm90", "glasberg_moore_90", "glasberg_moore")
@elementwise("freq", 0)
def erb(freq, Hz=None):
  """Equivalent rectangular bandwidth, Glasberg & Moore (1990): ``24.7 (4.37e-3 f + 1)``.
  ``freq`` in Hz, or in rad/sample when ``Hz = sHz(rate)[1]`` is given (the result is then
  in rad/sample too). Reference ``lazy_auditory.py:55-70``."""
  hertz, unit = _in_hertz(freq, Hz)
  return 24.7 * (4.37e-3 * hertz + 1.) * unit


@erb.strategy("mg83", "moore_glasberg_83")
@elementwise("freq", 0)
def erb(freq, Hz=None):
  """Equivalent rectangular bandwidth, Moore & Glasberg (1983):
  ``6.23e-6 f**2 + 93.39e-3 f + 28.52``. Reference ``lazy_auditory.py:73-88``."""
  hertz, unit = _in_hertz(freq, Hz)
  return (6.23e-6 * hertz ** 2 + 93.39e-3 * hertz + 28.52) * unit


def gammatone_erb_constants(n):
  """``(1/a_n, c_n)`` of Holdsworth et al. (1988) for an order-``n`` gammatone: the first
  compensates the ERB into the gammatone bandwidth parameter (1.019 for n = 4), the second
  gives the 3 dB bandwidth. Reference ``lazy_auditory.py:91-125``."""
  tnt = 2 * n - 2
  return (factorial(n - 1) ** 2 / (pi * factorial(tnt) * 2 ** -tnt),
          2 * (2 ** (1. / n) - 1) ** .5)


gammatone = StrategyDict("gammatone")


@gammatone.strategy("sampled")
def gammatone(freq, bandwidth, phase=0, eta=4):
  """Gammatone from the sampled impulse response ``n**(eta-1) exp(-bandwidth n)
  cos(freq n + phase)`` (Bellini 2013); reference ``lazy_auditory.py:151-182``.
  ``freq`` and ``bandwidth`` in rad/sample. The first section carries the whole numerator."""
  assert eta >= 1
  A = exp(-bandwidth)
  numerator = cos(phase) - A * cos(freq - phase) * z ** -1
  denominator = 1 - 2 * A * cos(freq) * z ** -1 + A ** 2 * z ** -2
  filt = (numerator / denominator).diff(n=eta - 1, mul_after=-z)
  # the differentiated denominator lost precision: rebuild it from the single section
  f0 = ZFilter(filt.numpoly) / denominator
  f0 /= abs(f0.freq_response(freq))   # max gain == 1.0 (0 dB)
  fn = 1 / denominator
  fn /= abs(fn.freq_response(freq))
  return CascadeFilter([f0] + [fn] * (eta - 1))


@gammatone.strategy("slaney")
def gammatone(freq, bandwidth):
  """Slaney's (1993) cascade of four one-zero two-pole sections; reference
  ``lazy_auditory.py:185-202``."""
  radius = exp(-bandwidth)
  c, s = cos(freq), sin(freq)
  poles = 1 - 2 * radius * c * z ** -1 + radius ** 2 * z ** -2        # shared by the four sections
  sections = []
  for outer in (1., -1.):
    for inner in (1., -1.):
      zero = c + outer * (sqrt(2) + inner) * s                        # cos w +- (sqrt 2 +- 1) sin w
      section = (1 - radius * zero * z ** -1) / poles
      sections.append(section / abs(section.freq_response(freq)))     # 0 dB at the centre frequency
  return CascadeFilter(sections)


@gammatone.strategy("klapuri")
def gammatone(freq, bandwidth):
  """Klapuri's (2008) cascade of resonators ``[z_exp, poles_exp] * 2`` with doubled
  bandwidth; reference ``lazy_auditory.py:205-218``."""
  bw2 = bandwidth * 2
  resons = [resonator.z_exp, resonator.poles_exp] * 2
  return CascadeFilter(reson(freq, bw2) for reson in resons)


# --------------------------------------------------------------------------------------
# the ERB bank of the benchmark configurations (not a reference object)
# --------------------------------------------------------------------------------------
def erb_space(low=50.0, high=20000.0, n=64):
  """``n`` centre frequencies (Hz) equally spaced on the Glasberg-Moore ERB-rate scale
  ``E(f) = 21.4 log10(1 + 0.00437 f)`` -- the scale whose derivative is ``erb.gm90``."""
  E = lambda f: 21.4 * log10(1 + 0.00437 * f)
  Einv = lambda v: (10 ** (v / 21.4) - 1) / 0.00437
  lo, hi = E(low), E(high)
  if n == 1:
    return [float(low)]
  return [float(Einv(lo + i * (hi - lo) / (n - 1))) for i in range(n)]


def gammatone_bank(freqs=None, rate=48000, strategy="slaney", order=4, erb_model="gm90"):
  """Bank of gammatone cascades, one channel per centre frequency (Hz), each designed
  exactly as ``examples/gammatone_plots.py:47,64`` of the reference does:
  ``bw = gammatone_erb_constants(order)[0] * erb(fc * Hz, Hz)``; ``gammatone[strategy](fc * Hz, bw)``.
  Returns a :class:`~audiolazy_b200.bank.FilterBank`."""
  from .bank import FilterBank
  if freqs is None:
    freqs = erb_space()
  _, Hz = sHz(rate)
  design = gammatone[strategy] if isinstance(strategy, str) else strategy
  channels = []
  for fc in freqs:
    bw = gammatone_erb_constants(order)[0] * erb[erb_model](fc * Hz, Hz)
    channels.append(design(fc * Hz, bw))
  bank = FilterBank(channels)
  bank.freqs = list(freqs)
  bank.rate = rate
  return bank
