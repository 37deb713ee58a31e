"""Minimal lazy sequence container: the result type of every filter call.

The drop-in boundary of the hot path is "a filter is any callable that receives an
iterable as input and returns a Stream" (reference ``audiolazy/lazy_filters.py:975-978``).
This module re-creates the subset of ``Stream`` / ``thub`` (reference
``audiolazy/lazy_stream.py:74-405`` and ``:469-630``) that the filters, their tests and
their callers rely on: iteration, ``take/peek/skip/limit/copy/map/filter/append/blocks``,
elementwise operators and the tee hub used to fan one input out to several filters.
``ControlStream`` and ``Streamix`` (interactive mixing) are out of scope.
"""
from __future__ import annotations

import itertools as it
import math
import operator
from collections import deque
from collections.abc import Iterable

__all__ = ["Stream", "StreamTeeHub", "thub", "tostream", "avoid_stream"]


class Stream(Iterable):
  """Iterable with elementwise operators (a lazy 1-D array).

  ``Stream(iterable)`` wraps it; ``Stream(a, b, c)`` of non-iterables cycles over the
  values endlessly; ``Stream(it1, it2)`` chains; ``Stream(5)`` repeats 5 forever
  (reference ``lazy_stream.py:137-191``). Not thread-safe, single-use.
  """
  __slots__ = ("_data",)
  _ignored = ()   # classes whose operators take precedence (see avoid_stream)

  def __init__(self, *items):
    if not items:
      raise TypeError("Missing argument(s)")
    flags = [isinstance(i, Iterable) for i in items]
    if len(items) == 1:
      self._data = iter(items[0]) if flags[0] else it.repeat(items[0])
    elif all(flags):
      self._data = it.chain(*items)
    elif not any(flags):
      self._data = it.cycle(items)
    else:
      raise TypeError("Input with both iterables and non-iterables")

  def __iter__(self):
    return self._data

  def __bool__(self):
    raise TypeError("Streams can't be used as booleans.\n"
                    "Freeze it first (list(stream)) or use the bitwise operators &, | and ~.")

  # -- consuming helpers -------------------------------------------------------------
  def take(self, n=None, constructor=list):
    """First ``n`` items (fewer if the stream ends); ``take()`` gives one bare item."""
    if n is None:
      return next(self._data)
    if isinstance(n, float):
      if math.isinf(n) and n > 0:
        return constructor(self._data)
      n = int(round(n)) if n > 0 else 0   # nan and -inf give nothing
    return constructor(it.islice(self._data, max(int(n), 0)))

  def copy(self):
    """Tee: keeps this stream usable and returns an independent copy."""
    self._data, other = it.tee(self._data)
    return Stream(other)

  tee = copy

  def peek(self, n=None, constructor=list):
    """Like :meth:`take` without consuming."""
    return self.copy().take(n=n, constructor=constructor)

  def skip(self, n):
    """Lazily throw the first ``n`` items away."""
    data, count = self._data, int(round(n))

    def skipped():
      for _ in it.islice(data, count):
        pass
      for item in data:
        yield item

    self._data = skipped()
    return self

  def limit(self, n):
    """End the stream after ``n`` items."""
    self._data = it.islice(self._data, int(round(n)))
    return self

  def append(self, *other):
    self._data = it.chain(self._data, iter(Stream(*other)))
    return self

  def map(self, func):
    self._data = map(func, self._data)
    return self

  def filter(self, func):
    self._data = filter(func, self._data)
    return self

  def blocks(self, size, hop=None, padval=0.0):
    """Stream of lists of ``size`` items; each block starts ``hop`` items after the
    previous one and the last one is padded with ``padval`` when incomplete
    (semantics of reference ``lazy_misc.py:74-129``, with lists instead of deques)."""
    hop = size if hop is None else hop
    if size < 1 or hop < 1:
      raise ValueError("size and hop must be positive")
    src = self._data

    def gen():
      buf = list(it.islice(src, size))
      while len(buf) == size:
        yield list(buf)
        if hop >= size:
          for _ in it.islice(src, hop - size):
            pass
          buf = list(it.islice(src, size))
        else:
          fresh = list(it.islice(src, hop))
          buf = buf[hop:] + fresh
          if len(fresh) < hop:
            # buf now holds the kept overlap plus the new items: a partial block
            # exists only if at least one new item arrived
            if not fresh:
              return
            break
      if buf and (hop >= size or len(buf) > size - hop):
        yield buf + [padval] * (size - len(buf))

    return Stream(gen())

  def __abs__(self):
    return Stream(map(abs, self._data))

  def __getattr__(self, name):
    if name in ("__next__", "next"):
      raise AttributeError("Streams are iterable, not iterators")
    if name.startswith("__"):
      raise AttributeError(name)
    return Stream(getattr(item, name) for item in self._data)

  def __call__(self, *args, **kwargs):
    return Stream(item(*args, **kwargs) for item in self._data)

  @classmethod
  def register_ignored_class(cls, ignored):
    Stream._ignored = Stream._ignored + (ignored,)


def _binary(func, reverse=False):
  def method(self, other):
    if isinstance(other, Stream._ignored):
      return NotImplemented
    if isinstance(other, Iterable):
      pairs = (iter(other), iter(self)) if reverse else (iter(self), iter(other))
      return Stream(map(func, *pairs))
    if reverse:
      return Stream(map(lambda item: func(other, item), iter(self)))
    return Stream(map(lambda item: func(item, other), iter(self)))
  return method


for _name, _func in [("add", operator.add), ("sub", operator.sub), ("mul", operator.mul),
                     ("truediv", operator.truediv), ("floordiv", operator.floordiv), ("mod", operator.mod),
                     ("pow", operator.pow), ("and", operator.and_), ("or", operator.or_), ("xor", operator.xor),
                     ("lshift", operator.lshift), ("rshift", operator.rshift)]:
  setattr(Stream, "__%s__" % _name, _binary(_func))
  setattr(Stream, "__r%s__" % _name, _binary(_func, reverse=True))
for _name, _func in [("lt", operator.lt), ("le", operator.le), ("gt", operator.gt), ("ge", operator.ge),
                     ("eq", operator.eq), ("ne", operator.ne)]:
  setattr(Stream, "__%s__" % _name, _binary(_func))
Stream.__hash__ = object.__hash__
for _name, _func in [("neg", operator.neg), ("pos", operator.pos), ("invert", operator.invert)]:
  setattr(Stream, "__%s__" % _name, (lambda f: lambda self: Stream(map(f, iter(self))))(_func))
del _name, _func


def avoid_stream(cls):
  """Class decorator: ``stream <op> instance`` defers to the instance's reflected
  operator instead of iterating over it (filters are not sample sources)."""
  Stream.register_ignored_class(cls)
  return cls


def tostream(func):
  """Decorator turning a generator function into a Stream factory."""
  import functools

  @functools.wraps(func)
  def wrapper(*args, **kwargs):
    return Stream(func(*args, **kwargs))
  return wrapper


class StreamTeeHub(Stream):
  """A Stream that can be iterated a fixed number of times (``itertools.tee`` fan-out);
  each ``iter()`` / operator use consumes one copy (reference ``lazy_stream.py:469-571``)."""
  __slots__ = ("_copies",)

  def __init__(self, data, n):
    self._copies = deque(it.tee(iter(data), int(n)))
    self._data = None

  def __iter__(self):
    try:
      return self._copies.popleft()
    except IndexError:
      raise IndexError("StreamTeeHub has no more copies left to use")

  def copy(self):
    if not self._copies:
      raise IndexError("StreamTeeHub has no more copies left to use")
    first, extra = it.tee(self._copies.popleft())
    self._copies.appendleft(first)
    return Stream(extra)

  def take(self, *args, **kwargs):
    return Stream(iter(self)).take(*args, **kwargs)

  def peek(self, *args, **kwargs):
    return self.copy().take(*args, **kwargs)

  def skip(self, n):
    return Stream(iter(self)).skip(n)

  def limit(self, n):
    return Stream(iter(self)).limit(n)

  def map(self, func):
    return Stream(iter(self)).map(func)

  def filter(self, func):
    return Stream(iter(self)).filter(func)

  def append(self, *other):
    return Stream(iter(self)).append(*other)

  def __abs__(self):
    return Stream(map(abs, iter(self)))


def thub(data, n):
  """Tee hub: lets ``data`` be used ``n`` times in an expression. Non-iterables (plain
  numbers) are returned unchanged (reference ``lazy_stream.py:573-630``)."""
  return StreamTeeHub(data, n) if isinstance(data, Iterable) else data
