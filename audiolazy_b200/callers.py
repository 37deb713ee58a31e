"""Callers of the hot path that are pure filter applications.

Each is the reference's own one-liner around a linear filter, so each now runs the
filtering on the GPU through :class:`~audiolazy_b200.filters.ZFilter`:

* ``envelope.rms / .abs / .squared`` -- reference ``audiolazy/lazy_analysis.py:440-520``
* ``maverage.recursive / .fir / .deque`` -- ``lazy_analysis.py:523-616``
* ``karplus_strong`` -- ``audiolazy/lazy_synth.py:624-657`` (feedback comb with seeded memory)
* ``accumulate_z`` -- ``audiolazy/lazy_itertools.py:82`` (``1 / (1 - z**-1)``)
* the tiny sources those callers and the tests need: ``zeros``, ``ones``, ``impulse``,
  ``white_noise`` (``lazy_synth.py:394-415``, ``:597-621``).
"""
from __future__ import annotations

import itertools as it
import random
from collections import deque
from math import isinf, pi

from .core import StrategyDict
from .filters import comb, lowpass, z
from .stream import Stream, thub, tostream

__all__ = ["envelope", "maverage", "karplus_strong", "accumulate_z", "zeros", "ones", "impulse", "white_noise"]


# ---- sources -----------------------------------------------------------------------------
def _endless(dur):
  return dur is None or (isinf(dur) and dur > 0)


@tostream
def zeros(dur=None, zero=0.):
  """``zero`` repeated ``dur`` times (endless if None)."""
  return it.repeat(zero) if _endless(dur) else it.repeat(zero, int(round(dur)))


@tostream
def ones(dur=None, one=1.):
  return it.repeat(one) if _endless(dur) else it.repeat(one, int(round(dur)))


@tostream
def impulse(dur=None, one=1., zero=0.):
  """``one`` followed by ``zero``s (reference ``lazy_synth.py:597-621``)."""
  if _endless(dur):
    return it.chain([one], it.repeat(zero))
  if dur >= .5:
    return it.chain([one], it.repeat(zero, int(dur - .5)))
  return iter(())


@tostream
def white_noise(dur=None, low=-1., high=1.):
  """Uniform random samples (reference ``lazy_synth.py:394-415``)."""
  if _endless(dur):
    return (random.uniform(low, high) for _ in it.count())
  return (random.uniform(low, high) for _ in range(int(round(dur))))


# ---- envelope ----------------------------------------------------------------------------
envelope = StrategyDict("envelope")


def _envelope(sig, cutoff, pre, post):
  """``post(lowpass(cutoff)(pre(sig)))`` (reference ``lazy_analysis.py:440-520``). With a constant cutoff the
  rectifier / squarer and the square root run on whole blocks next to the device launch (numpy, float64, the values
  the reference's Stream arithmetic produces) instead of one Python frame per sample; a Stream-valued cutoff keeps
  the elementwise Stream expression around the time-varying filter."""
  import numpy as np
  from . import _engine
  from .filters import _seed_histories
  filt = lowpass(cutoff)
  ops = {"square": (np.square, lambda s: s ** 2), "abs": (np.abs, abs), "sqrt": (np.sqrt, lambda s: s ** .5), None: (None, lambda s: s)}
  if filt.is_lti():
    secs = filt.sections()
    xinit, yinit = _seed_histories(secs, None, 0.)
    return _engine.filter_stream([secs], sig, [xinit], [yinit], pre=ops[pre][0], post=ops[post][0])
  return ops[post][1](filt(ops[pre][1](thub(sig, 1))))


@envelope.strategy("rms")
def envelope(sig, cutoff=pi / 512):
  """RMS envelope: lowpass of the squared signal, then square root."""
  return _envelope(sig, cutoff, "square", "sqrt")


@envelope.strategy("abs")
def envelope(sig, cutoff=pi / 512):
  """Lowpass of the rectified signal."""
  return _envelope(sig, cutoff, "abs", None)


@envelope.strategy("squared")
def envelope(sig, cutoff=pi / 512):
  """Lowpass of the squared signal."""
  return _envelope(sig, cutoff, "square", None)


# ---- moving average ----------------------------------------------------------------------
maverage = StrategyDict("maverage")


@maverage.strategy("deque")
def maverage(size):
  """Running mean with a deque (the reference's only non-ZFilter strategy; host-side)."""
  scale = 1. / size

  @tostream
  def running_mean(sig, zero=0.):
    # the window holds the already scaled samples; the running total drops the oldest and adds
    # the newest (same operation order as the reference, so results are bit-identical)
    window = deque([zero * scale] * size)
    total = zero
    for sample in sig:
      total -= window.popleft()
      scaled = sample * scale
      window.append(scaled)
      total += scaled
      yield total

  return running_mean


@maverage.strategy("recursive", "feedback")
def maverage(size):
  """``(1/size) (1 - z**-size) / (1 - z**-1)`` as a ZFilter."""
  return (1. / size) * (1 - z ** -size) / (1 - z ** -1)


@maverage.strategy("fir")
def maverage(size):
  """``sum((1/size) z**-i)`` as a FIR ZFilter."""
  return sum((1. / size) * z ** -i for i in range(size))


# ---- Karplus-Strong ----------------------------------------------------------------------
def karplus_strong(freq, tau=2e4, memory=white_noise):
  """"Digitar" synthesis: a feedback comb whose delay line starts filled with ``memory``
  (a callable receiving the size, or an iterable), fed with silence."""
  return comb.tau(2 * pi / freq, tau).linearize()(zeros(), memory=memory)


accumulate_z = 1 / (1 - z ** -1)
