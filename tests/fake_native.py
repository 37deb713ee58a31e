"""A stand-in for the native layer so that the HOST logic (filter objects -> section tables ->
block pump -> lazy Streams) can be exercised without a GPU.

Test infrastructure only.  ``install(monkeypatch)`` replaces ``audiolazy_b200._capi.Plan`` with a
plan whose arithmetic is the CPU oracle (float64, then rounded to float32 as the device stores it)
and gives ``audiolazy_b200._engine`` a torch shim whose "cuda" device is the CPU.  Nothing of this
is reachable from the product: without ``install`` a filter call on a box without a GPU raises.
"""
import ctypes
import types

import numpy as np

import oracle


def _f32(ptr, n):
  return np.ctypeslib.as_array((ctypes.c_float * int(n)).from_address(int(ptr)))


def _f64(ptr, n):
  return np.ctypeslib.as_array((ctypes.c_double * int(n)).from_address(int(ptr)))


class FakePlan(object):
  """Same surface as ``_capi.Plan`` for what the engine uses."""
  states = {}      # state pointer -> dict

  def __init__(self, bank, force_generic=False, parallel=False, **_flags):
    self.bank = [[([float(v) for v in b] or [0.0], [float(v) for v in a]) for b, a in ch] for ch in bank]
    for ch in self.bank:
      for b, a in ch:
        if a[0] == 0:
          from audiolazy_b200._capi import NativeError
          raise NativeError("Invalid filter gain (a0 == 0)", -3)
    self.n_channels = len(self.bank)
    self.n_sections = max([len(ch) for ch in self.bank] + [1])
    self.xd = max([len(b) - 1 for ch in self.bank for b, _ in ch] + [0])
    self.yd = max([len(a) - 1 for ch in self.bank for _, a in ch] + [0])
    self.kind = 2 if force_generic else 1
    self.num_taps = max([len(b) for ch in self.bank for b, _ in ch] + [1])
    self.launches = 0

  # ---- state: the whole input history is kept and the oracle re-runs from the start (tests are small)
  def state_doubles(self, n_streams):
    return max(1, int(n_streams))

  def state_init(self, state_ptr, n_streams, xinit=None, yinit=None, stream=0):
    pad = lambda arr, depth: None if arr is None else np.asarray(arr, dtype=np.float64).reshape(
      self.n_channels, self.n_sections, -1)
    FakePlan.states[int(state_ptr)] = {"x": np.zeros((int(n_streams), 0), dtype=np.float32),
                                       "xi": pad(xinit, self.xd), "yi": pad(yinit, self.yd), "tv": None}

  def apply(self, x_ptr, y_ptr, state_ptr, n_streams, n_samples, x_stride, y_stride, stream=0):
    S, T, C = int(n_streams), int(n_samples), self.n_channels
    st = FakePlan.states[int(state_ptr)]
    x = _f32(x_ptr, (S - 1) * x_stride + T).copy() if S > 1 else _f32(x_ptr, T).copy()
    rows = np.stack([x[s * x_stride:s * x_stride + T] for s in range(S)])
    st["x"] = np.concatenate([st["x"], rows], axis=1)
    full = oracle.bank_apply(st["x"], self._padded_bank(), xinit=st["xi"], yinit=st["yi"])
    y = _f32(y_ptr, (S * C - 1) * y_stride + T)
    for s in range(S):
      for c in range(C):
        off = (s * C + c) * y_stride
        y[off:off + T] = full[s, c, -T:].astype(np.float32)
    self.launches += 1

  def apply_sum(self, x_ptr, out_ptr, state_ptr, n_streams, n_samples, x_stride, out_stride, stream=0):
    """ParallelFilter in one call: float64 channel results summed left to right, rounded to float32 once."""
    S, T = int(n_streams), int(n_samples)
    st = FakePlan.states[int(state_ptr)]
    x = _f32(x_ptr, (S - 1) * x_stride + T).copy() if S > 1 else _f32(x_ptr, T).copy()
    rows = np.stack([x[s * x_stride:s * x_stride + T] for s in range(S)])
    st["x"] = np.concatenate([st["x"], rows], axis=1)
    full = oracle.bank_apply(st["x"], self._padded_bank(), xinit=st["xi"], yinit=st["yi"])
    out = _f32(out_ptr, (S - 1) * out_stride + T)
    for s in range(S):
      acc = full[s, 0, -T:].copy()
      for c in range(1, self.n_channels):
        acc = acc + full[s, c, -T:]
      out[s * out_stride:s * out_stride + T] = acc.astype(np.float32)
    self.launches += 1

  def _padded_bank(self):
    # the oracle wants every channel to have the same number of sections: absent = identity
    return [ch + [([1.0], [1.0])] * (self.n_sections - len(ch)) for ch in self.bank]

  # ---- time-varying single filter (generic plan): direct evaluation of the difference equation
  def taps(self):
    (b, a), = self.bank[0]
    return [(k, False) for k, v in enumerate(b) if v != 0] + [(k, True) for k, v in enumerate(a) if k >= 1 and v != 0]

  def apply_tv(self, x_ptr, y_ptr, state_ptr, n_streams, n_samples, x_stride, y_stride, coef_ptr, coef_stride, stream=0):
    assert int(n_streams) == 1
    T = int(n_samples)
    st = FakePlan.states[int(state_ptr)]
    if st["tv"] is None:
      xi = [] if st["xi"] is None else st["xi"][0, 0].tolist()
      yi = [] if st["yi"] is None else st["yi"][0, 0].tolist()
      st["tv"] = {"xh": xi + [0.0] * (self.xd - len(xi)), "yh": yi + [0.0] * (self.yd - len(yi))}
    xh, yh = st["tv"]["xh"], st["tv"]["yh"]      # xh[d-1] = x[n-d], yh[d-1] = y[n-d]
    taps = self.taps()
    coef = _f64(coef_ptr, (len(taps) - 1) * coef_stride + T) if taps else np.zeros(0)
    x, y = _f32(x_ptr, T), _f32(y_ptr, T)
    for n in range(T):
      xn = float(x[n])
      acc = 0.0
      for row, (delay, is_den) in enumerate(taps):
        value = coef[row * coef_stride + n]
        past = (yh[delay - 1] if is_den else (xn if delay == 0 else xh[delay - 1]))
        acc += value * past
      if xh:
        xh.insert(0, xn); xh.pop()
      if yh:
        yh.insert(0, acc); yh.pop()
      y[n] = np.float32(acc)
    self.launches += 1

  def apply_host(self, x, y=None, state_ptr=None):
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    out = oracle.bank_apply(x, self._padded_bank()).astype(np.float32)
    if y is not None:
      y[...] = out
      return y
    return out


def _sum_channels(y_ptr, out_ptr, n_streams, n_channels, n_samples, y_stride, out_stride, stream=0):
  S, C, T = int(n_streams), int(n_channels), int(n_samples)
  y = _f32(y_ptr, (S * C - 1) * y_stride + T)
  out = _f32(out_ptr, (S - 1) * out_stride + T)
  for s in range(S):
    acc = np.zeros(T, dtype=np.float64)
    for c in range(C):
      off = (s * C + c) * y_stride
      acc = acc + y[off:off + T].astype(np.float64)      # left associated, as ParallelFilter adds
    out[s * out_stride:s * out_stride + T] = acc.astype(np.float32)


class _TorchShim(object):
  """torch with a 'cuda' device that is the CPU (tensors are host tensors, pointers host pointers)."""

  def __init__(self):
    import torch
    self._torch = torch
    stream = types.SimpleNamespace(cuda_stream=0)
    self.cuda = types.SimpleNamespace(is_available=lambda: True, current_device=lambda: 0,
                                      current_stream=lambda *a, **k: stream, set_device=lambda *a: None,
                                      device_count=lambda: 1)

  def device(self, *args, **kwargs):
    return self._torch.device("cpu")

  def __getattr__(self, name):
    return getattr(self._torch, name)


def install(monkeypatch):
  from audiolazy_b200 import _capi, _engine
  shim = _TorchShim()
  monkeypatch.setattr(_capi, "Plan", FakePlan)
  monkeypatch.setattr(_capi, "set_device", lambda index: None)
  monkeypatch.setattr(_capi, "sum_channels", _sum_channels)
  monkeypatch.setattr(_engine, "torch_mod", lambda: shim)
  monkeypatch.setattr(_engine, "_cache", {})
  FakePlan.states.clear()
  return shim
