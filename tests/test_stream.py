"""The minimal Stream / thub / Poly API the filter path hands its results in."""
import itertools as it

import pytest

from audiolazy_b200 import Poly, Stream, thub, x, zero_pad
from audiolazy_b200.stream import StreamTeeHub


def test_constructors_and_operators():
  assert (Stream([1, 2, 3]) + Stream([8, 5])).take(10) == [9, 7]
  assert (Stream(1, 2, 3) + Stream(8, 5)).take(7) == [9, 7, 11, 6, 10, 8, 9]
  assert Stream(5).take(3) == [5, 5, 5]
  assert (2 * Stream(it.count()) + Stream(3)).take(4) == [3, 5, 7, 9]
  assert (1 - Stream([1, 2, 3])).take(3) == [0, -1, -2]
  assert (-Stream([1, -2])).take(2) == [-1, 2] and abs(Stream([1, -2])).take(2) == [1, 2]
  assert (Stream([1, 4]) ** .5).take(2) == [1.0, 2.0]
  assert (Stream([1, 2, 3]) > 1).take(3) == [False, True, True]
  with pytest.raises(TypeError):
    Stream()
  with pytest.raises(TypeError):
    Stream([1], 2)
  with pytest.raises(TypeError):
    bool(Stream([1]))


def test_take_peek_skip_limit_copy_map():
  s = Stream(it.count())
  assert s.take() == 0 and s.take(3) == [1, 2, 3] and s.peek(2) == [4, 5] and s.take(2) == [4, 5]
  assert Stream([1, 2]).take(5) == [1, 2]                    # short streams give what they have
  assert Stream([4, 3, 2, 3, 2]).take(3.6) == [4, 3, 2, 3]
  assert Stream([4, 3, 2]).take(float("inf")) == [4, 3, 2]
  assert Stream(range(10)).skip(7).take(5) == [7, 8, 9]
  assert list(Stream(it.count()).limit(3)) == [0, 1, 2]
  a = Stream([1, 2, 3])
  b = a.copy()
  assert list(a) == list(b) == [1, 2, 3]
  assert Stream([1, 2]).append([3], [4]).take(9) == [1, 2, 3, 4]
  assert Stream(it.count()).map(float).filter(lambda v: v % 2).take(2) == [1.0, 3.0]
  assert Stream(range(10)).blocks(4).take(3) == [[0, 1, 2, 3], [4, 5, 6, 7], [8, 9, 0.0, 0.0]]
  assert Stream(range(6)).blocks(4, hop=2).take(3) == [[0, 1, 2, 3], [2, 3, 4, 5]]
  with pytest.raises(StopIteration):
    Stream([]).take()


def test_thub():
  h = thub([1, 2, 3], 2)
  assert isinstance(h, StreamTeeHub)
  assert list(h + h) == [2, 4, 6]
  with pytest.raises(IndexError):
    iter(h)
  assert thub(3.5, 4) == 3.5
  assert list(zero_pad([1, 2], 2, 1, zero=9)) == [9, 9, 1, 2, 9]


def test_poly():
  p = x ** 5 - x + 7
  assert dict(p.terms()) == {0: 7, 1: -1, 5: 1} and p.order == 5 and len(p) == 3
  assert (x + 2)(17) == 19 and (x ** 2 + 2 * x + 1)(.5) == 2.25
  assert list((x ** 2 + 1).values()) == [1, 0.0, 1]
  assert Poly([1, 0, 2]) == Poly({0: 1, 2: 2}) and len(Poly([0, 0.0])) == 0
  assert (x ** -2 + x).is_laurent() and not (x ** -2 + x).is_polynomial()
  assert dict(((x + 1) * (x - 1)).terms()) == {0: -1, 2: 1}      # the x terms cancel and are dropped
  assert dict((x ** 3).diff().terms()) == {2: 3}
  assert dict((x + 1)(Poly({-1: 1})).terms()) == {-1: 1, 0: 1}
  assert sorted(Poly([-2, 1]).roots) == [2.0]
  with pytest.raises(AttributeError):
    (x ** -3 + 4).order
  with pytest.raises(NotImplementedError):
    (x + 1) / (x + 2)


def test_strategy_dict_dictionary_views():
  """keys / items / values / key2keys / value2keys as the reference's MultiKeyDict gives them:
  one entry per strategy, keyed by the tuple of all its names (lazy_core.py:431-659)."""
  from audiolazy_b200 import erb, lowpass, gammatone
  assert erb.keys() == [("gm90", "glasberg_moore_90", "glasberg_moore"), ("mg83", "moore_glasberg_83")]
  assert [(k, f.__name__) for k, f in erb.items()] == [(erb.keys()[0], "gm90"), (erb.keys()[1], "mg83")]
  assert erb.values() == [erb.gm90, erb.mg83] and list(erb) == erb.values()
  assert erb.key2keys("moore_glasberg_83") == ("mg83", "moore_glasberg_83") and erb.key2keys("nope") == ()
  assert erb.value2keys(erb.gm90) == erb.keys()[0]
  assert [k[0] for k in gammatone.keys()] == ["sampled", "slaney", "klapuri"]
  assert lowpass.default is lowpass.pole and len(lowpass) == 4
