"""Data formats either side of the path.

* :func:`chunks` -- pack a sample iterable into fixed-size binary blocks, float32 by default,
  2048 samples per block (reference ``audiolazy/lazy_io.py:44-128``: what its audio thread
  writes to the sound card).
* :class:`WavStream` -- PCM wave file (8/16/24/32-bit little endian) to a float Stream in
  ``[-1, 1)`` (reference ``audiolazy/lazy_wav.py:31-130``).
* :func:`wav_batch` / :func:`pcm_to_float32` -- the same decoding for whole files at once into
  the ``[streams][samples]`` float32 batches the device path consumes.
"""
from __future__ import annotations

import itertools as it
import struct
import wave

import numpy as np

from .core import StrategyDict
from .stream import Stream

__all__ = ["chunks", "WavStream", "wav_batch", "pcm_to_float32"]

chunks = StrategyDict("chunks")
DEFAULT_CHUNK = 2048   # samples (reference lazy_io.py:45)


@chunks.strategy("struct", "array")
def chunks(seq, size=None, dfmt="f", byte_order=None, padval=0.):
  """Generator of ``bytes`` blocks of ``size`` samples packed as ``dfmt`` (struct codes:
  ``f`` float32, ``d`` float64, ``h`` int16, ...); the last block is padded with ``padval``."""
  if size is None:
    size = DEFAULT_CHUNK
  packer = struct.Struct((byte_order or "") + str(size) + dfmt)
  src = iter(seq)
  while True:
    block = list(it.islice(src, size))
    if not block:
      return
    if len(block) < size:
      block += [padval] * (size - len(block))
    yield packer.pack(*block)


def pcm_to_float32(raw, bits, keep=False, dtype=np.float32):
  """Decode little-endian PCM bytes to samples in ``[-1, 1)`` (``value / 2**(bits-1)``; 8-bit data is unsigned,
  offset 128), or the stored integers when ``keep``. ``dtype=np.float32`` (default) is what the device path
  consumes; ``np.float64`` is exact for every width, as the reference's ``int / int`` (``lazy_wav.py:117-123``)."""
  if bits == 8:
    data = np.frombuffer(raw, dtype=np.uint8).astype(np.int32) - (0 if keep else 128)
  elif bits == 16:
    data = np.frombuffer(raw, dtype="<i2").astype(np.int32)
  elif bits == 24:
    b = np.frombuffer(raw, dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    data = (b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16))
    data = np.where(data >= (1 << 23), data - (1 << 24), data)
  elif bits == 32:
    data = np.frombuffer(raw, dtype="<i4").astype(np.int64)
  else:
    raise ValueError("unsupported sample width: %d bits" % bits)
  if keep:
    return data
  return (data / float(1 << (bits - 1))).astype(dtype)


class WavStream(Stream):
  """Stream of the samples of a PCM wave file, scaled to ``[-1, 1)`` (or raw ints with
  ``keep=True``); stereo data is interleaved. Attributes: ``rate``, ``channels``, ``bits``."""
  __slots__ = ("rate", "channels", "bits", "_file")

  def __init__(self, wave_file, keep=False):
    self._file = wave.open(wave_file, "rb")
    self.rate = self._file.getframerate()
    self.channels = self._file.getnchannels()
    self.bits = 8 * self._file.getsampwidth()

    def data():
      w = self._file
      try:
        while True:
          raw = w.readframes(4096)
          if not raw:
            break
          # float64: 24- and 32-bit samples do not fit a float32 mantissa, the reference yields int / 2**(bits-1) exactly
          for value in pcm_to_float32(raw, self.bits, keep=keep, dtype=np.float64).tolist():
            yield value
      finally:
        w.close()

    Stream.__init__(self, data())


def wav_batch(paths, channel=0):
  """Load mono data (one channel of each file) as a float32 batch ``[len(paths)][max_len]``,
  zero padded; returns ``(batch, lengths, rates)``."""
  rows, rates = [], []
  for path in paths:
    with wave.open(path, "rb") as w:
      bits, nch = 8 * w.getsampwidth(), w.getnchannels()
      data = pcm_to_float32(w.readframes(w.getnframes()), bits)
      rows.append(data[channel::nch])
      rates.append(w.getframerate())
  lengths = [len(r) for r in rows]
  batch = np.zeros((len(rows), max(lengths) if lengths else 0), dtype=np.float32)
  for i, r in enumerate(rows):
    batch[i, :len(r)] = r
  return batch, lengths, rates
