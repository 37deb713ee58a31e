"""The host layer (filter objects, section tables, seeding, block pump, lazy Streams, bank fan-out,
time-varying plumbing) without a GPU: the native layer is replaced by tests/fake_native.py, whose
arithmetic is the CPU oracle.  The same assertions run against the real kernels in
tests/test_gpu_api.py; here they guard the Python side in the CPU suite."""
import itertools as it

import numpy as np
import pytest

import audiolazy_b200 as ab
import fake_native
import test_gpu_api as gpu_cases
from conftest import rel_err, signal


@pytest.fixture
def fake(monkeypatch):
  fake_native.install(monkeypatch)
  return ab


def test_reference_doctests(fake):
  gpu_cases.test_reference_doctests(fake)


def test_lfilter_grid(fake):
  gpu_cases.test_lfilter_grid(fake)


def test_identity_gain_delay_empty_lists(fake):
  gpu_cases.test_identity_gain_delay_empty_lists(fake)


def test_callers_of_the_path(fake, vectors):
  gpu_cases.test_callers_of_the_path(fake, vectors)


def test_time_varying_coefficients(fake, vectors):
  gpu_cases.test_time_varying_coefficients(fake, vectors)


def test_stream_valued_design_parameters_filter(fake, vectors):
  gpu_cases.test_stream_valued_design_parameters_filter(fake, vectors)


def test_seeding_and_bank_fan_out(fake, designs, vectors):
  """memory= / zero= reach the device state in the right slots; FilterBank(seq) returns one lazy
  Stream per channel fed by one pump; blocks grow geometrically and state is carried between them."""
  xs = signal(3, 64).tolist()
  f = ab.ZFilter([0.5, -0.25, 2.0], [2.0, 0.5, -0.3])
  assert rel_err(list(f(xs, memory=[0.75, -1.5], zero=0.125)), vectors["seed_single_y"]) <= 1e-6
  assert rel_err(list(f(xs, memory=[0.75], zero=-0.5)), vectors["seed_short_memory_y"]) <= 1e-6
  bank = ab.FilterBank([ab.gammatone.slaney(0.3, 0.05), ab.gammatone.klapuri(0.6, 0.04), 1 / (1 - 0.5 * ab.z ** -1)])
  x = signal(21, 3000)
  streams = bank(iter(x.tolist()))             # an iterator: pulled block by block (256, 1024, ...)
  assert len(streams) == 3 and all(isinstance(s, ab.Stream) for s in streams)
  first = streams[2].take(10)                  # one channel far ahead of the others
  outs = [list(s) for s in streams]
  want = [np.array(list(ch(x.tolist()))) for ch in bank]
  assert rel_err(first + outs[2], want[2]) <= 1e-6
  assert rel_err(outs[0], want[0]) <= 1e-6 and rel_err(outs[1], want[1]) <= 1e-6
  # endless input, finite take
  acc = 1 / (1 - ab.z ** -1)
  assert acc(ab.Stream(it.count())).take(6) == [0.0, 1.0, 3.0, 6.0, 10.0, 15.0]


def test_errors_raise_at_call_time(fake):
  with pytest.raises(ValueError, match="Non-causal filter"):
    (ab.z + 1)([1.0, 2.0])
  broken = ab.ZFilter([1.0], [1.0, 0.5])
  broken.denpoly = ab.Poly({1: 0.5})           # a0 == 0 is only reachable by hand (the constructor normalises)
  with pytest.raises(ZeroDivisionError, match="Invalid filter gain"):
    broken([1.0])
  with pytest.raises(NotImplementedError):
    ab.ZFilter([1j, 1.0])([1.0, 2.0])           # complex coefficients have no accelerated path (and no fallback)


def test_configs_through_the_python_api(fake, designs, vectors):
  gpu_cases.test_configs_through_the_python_api(fake, designs, vectors)


@pytest.mark.parametrize("strategy", ["slaney", "sampled"])
def test_gammatone_channels_and_bank(fake, vectors, strategy):
  gpu_cases.test_gammatone_channels_and_bank(fake, vectors, strategy)
