"""Filterbank object: one input fanned out to many filters.

The reference has no bank class -- its only bank usage is a Python loop over centre
frequencies applying each cascade to a copy of the input
(``examples/gammatone_plots.py:42-73``; fan-out through ``thub``, reference
``audiolazy/lazy_stream.py:573-630``). :class:`FilterBank` is that loop as an object,
with each channel exactly the reference's filter, evaluated for all channels (and any
number of independent input streams) by ONE kernel launch per block.
"""
from __future__ import annotations

import numpy as np

from . import _engine
from .filters import CascadeFilter, FilterList, LinearFilter, _seed_histories

__all__ = ["FilterBank", "BankState"]


class BankState(object):
  """Device state of a bank for ``n_streams`` input streams: carries every recurrence
  from one :meth:`FilterBank.apply` call to the next (endless inputs in blocks)."""

  def __init__(self, bank, n_streams, memory=None, zero=0.):
    self.bank = bank
    self.n_streams = int(n_streams)
    seeds = [_seed_histories(ch, memory, zero) for ch in bank.sections()]
    self.tensor = bank.device_bank().new_state(self.n_streams, [s[0] for s in seeds], [s[1] for s in seeds])


class FilterBank(list):
  """List of LTI filters (``ZFilter`` / all-LTI ``CascadeFilter``) sharing one input.

  * ``bank(seq)`` -> list of Streams, one per channel (``[f(seq_copy) for f in bank]``).
  * ``bank.apply(x)`` -> CUDA tensor ``y[S, C, T]`` for a CUDA float32 tensor ``x[S, T]``
    of ``S`` independent streams (``channel_major=True``: ``y[C, S, T]``); pass
    ``state=bank.new_state(S)`` to continue streams across calls.
  * ``bank.apply_host(x)`` -> the same through numpy host buffers (copies inside).
  """

  def __init__(self, filters=()):
    list.__init__(self, filters)
    self._sections = None
    self._sections_of = None
    self.freqs = None
    self.rate = None

  def sections(self):
    """Channels -> sections -> ``(b, a)``: the table handed to ``alz_plan_create``."""
    ids = tuple(id(f) for f in self)                   # the cache follows any mutation of the list
    if self._sections is None or self._sections_of != ids:
      self._sections_of = ids
      table = []
      for f in self:
        if isinstance(f, CascadeFilter):
          secs = f._flat_sections()
          if secs is None:
            raise NotImplementedError("bank channels must be LTI filters")
        elif isinstance(f, LinearFilter):
          secs = f.sections()
        elif not callable(f):
          secs = LinearFilter(f).sections()
        else:
          raise NotImplementedError("bank channels must be LTI filters")
        table.append(secs)
      self._sections = table
    return self._sections

  def device_bank(self):
    return _engine.device_bank(self.sections())

  def new_state(self, n_streams, memory=None, zero=0.):
    return BankState(self, n_streams, memory=memory, zero=zero)

  # -- lazy API ------------------------------------------------------------------------
  def __call__(self, seq, memory=None, zero=0.):
    secs = self.sections()
    seeds = [_seed_histories(ch, memory, zero) for ch in secs]
    return _engine.bank_streams(secs, seq, [s[0] for s in seeds], [s[1] for s in seeds])

  # -- batch API -----------------------------------------------------------------------
  def apply(self, x, state=None, out=None, channel_major=False):
    db = self.device_bank()
    if x.dim() == 1:
      x = x.unsqueeze(0)
    if state is None:
      state = self.new_state(x.shape[0])
    self._check_state(state, x.shape[0], db)
    return db.apply(x, state.tensor, out=out, channel_major=channel_major)

  def _check_state(self, state, n_streams, db):
    if state.n_streams != n_streams:
      raise ValueError("state was created for %d streams, x has %d" % (state.n_streams, n_streams))
    if state.tensor.numel() != max(1, db.plan.state_doubles(n_streams)) or state.tensor.device != db.device:
      raise ValueError("state belongs to another bank or device")

  def freq_response(self, freqs):
    """Complex128 ndarray ``[C, n]``: every channel's response on the grid ``freqs`` (rad/sample),
    evaluated on the device (the per-filter ``freq_response`` of reference
    ``lazy_filters.py:267-301``, batched over the bank)."""
    return self.device_bank().freq_response(np.asarray(freqs, dtype=np.float64)).cpu().numpy()

  # -- fused consumer --------------------------------------------------------------------
  @staticmethod
  def _envelope_pole(cutoff):
    from .filters import lowpass
    (b, a), = lowpass(cutoff).sections()      # the reference's envelope lowpass (lazy_analysis.py:440-520, lowpass.pole)
    if len(b) != 1 or len(a) != 2:
      raise NotImplementedError("the fused envelope uses a one-pole lowpass")
    return b[0] / a[0], -a[1] / a[0]

  def envelope(self, x, cutoff=np.pi / 512, decim=48, mode="abs"):
    """Channel envelopes of a CUDA float32 batch ``x[S, T]`` -> ``[S, C, T // decim]``: ``envelope.<mode>`` (abs /
    squared / rms, reference ``lazy_analysis.py:440-520``) of every channel output, decimated by ``decim`` -- rectifier,
    lowpass and decimation run inside the bank kernel, the channel signals never reach memory."""
    torch = _engine.torch_mod()
    db = self.device_bank()
    S, T = x.shape
    g, R = self._envelope_pole(cutoff)
    x = x.contiguous()
    env = torch.empty((S, len(self), T // decim), dtype=torch.float32, device=x.device)
    state = torch.zeros(max(1, db.plan.state_doubles(S)), dtype=torch.float64, device=x.device)
    env_state = torch.zeros(S * len(self), dtype=torch.float64, device=x.device)
    db.plan.apply_envelope(x.data_ptr(), env.data_ptr(), state.data_ptr(), env_state.data_ptr(), S, T, T, T // decim, decim,
                           mode, g, R, torch.cuda.current_stream(x.device).cuda_stream)
    return env

  def envelope_host(self, x, cutoff=np.pi / 512, decim=48, mode="abs", out=None):
    """:meth:`envelope` through host buffers (``alz_apply_envelope_f32_host``): a host caller receives ``256 / decim``
    bytes per input sample instead of the bank's 256."""
    g, R = self._envelope_pole(cutoff)
    return self.device_bank().plan.apply_envelope_host(x, out, decim=decim, mode=mode, g=g, R=R)

  def apply_host(self, x, out=None, state=None):
    """``x``: float32 ndarray ``[S, T]`` (or ``[T]``) on the host; returns ndarray ``[S, C, T]``.
    Goes through ``alz_apply_f32_host`` (pipelined H2D / kernel / D2H)."""
    db = self.device_bank()
    x = np.asarray(x, dtype=np.float32)
    state_ptr = None
    if state is not None:
      self._check_state(state, 1 if x.ndim == 1 else x.shape[0], db)
      # alz_apply_f32_host runs on private streams ordered after the legacy default stream only: whatever
      # produced the state on torch's current stream (new_state, a previous apply) must be complete
      _engine.torch_mod().cuda.current_stream(db.device).synchronize()
      state_ptr = state.tensor.data_ptr()
    return db.plan.apply_host(x, out, state_ptr)
