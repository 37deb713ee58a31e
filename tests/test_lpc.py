"""LPC analysis (reference audiolazy/lazy_lpc.py, lazy_analysis.py:277-342) against values the
reference itself produced (tests/golden/make_golden.py) and its doctests."""
import numpy as np
import pytest

import audiolazy_b200 as ab
from audiolazy_b200 import (ParCorError, ZFilter, acorr, lag_matrix, levinson_durbin, lpc, lsf, lsf_stable, parcor,
                            parcor_stable, toeplitz, z)
from conftest import rel_err, signal

RTOL = 1e-9


def close(got, want):
  got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
  return got.shape == want.shape and np.abs(got - want).max() <= RTOL * max(np.abs(want).max(), 1e-300)


def test_statistics_doctests():
  seq = [1, 2, 3, 4, 3, 4, 2]                                   # lazy_analysis.py:298-306
  assert acorr(seq) == [59, 52, 42, 30, 17, 8, 2]
  assert acorr(seq, 9) == [59, 52, 42, 30, 17, 8, 2, 0, 0, 0]
  assert acorr(seq, 3) == [59, 52, 42, 30]
  assert toeplitz([1, 2, 3]) == [[1, 2, 3], [2, 1, 2], [3, 2, 1]]
  assert lag_matrix([1, 2, 3, 4], 1) == [[29, 20], [20, 14]]
  with pytest.raises(ValueError):
    lag_matrix([1, 2, 3], 3)
  ld = levinson_durbin([12, 6, 0, -3, -6, -3, 0, 2, 4, 2], 3)   # lazy_lpc.py:93-99
  assert ld.numerator == [1, -0.625, 0.25, 0.125] and ld.error == 7.875
  filt = lpc.kautocor([-1, 0, 1, 0] * 4, 2)                     # lazy_lpc.py:244-254
  assert filt.numerator == [1, 0.0, 0.875] and filt.error == 1.875
  assert isinstance(filt, ZFilter) and str(filt) == "1 + 0.875 * z^-2"


def test_strategies_against_reference(designs, vectors):
  blk = vectors["lpc_blk"].astype(np.float64).tolist()
  gold = designs["lpc"]
  assert close(acorr(blk, 9), gold["acorr9"])
  assert close(lag_matrix(blk, 3), gold["lag_matrix3"])
  for case in gold["cases"]:
    filt = lpc[case["strategy"]](blk, case["order"])
    assert close(filt.numerator, case["numerator"]), case["strategy"]
    assert abs(filt.error - case["error"]) <= 1e-8 * abs(case["error"]), case["strategy"]
    assert filt.denominator == [1]
  assert lpc.acorr is lpc.autocor and lpc.cov is lpc.covar and lpc.kcov is lpc.kcovar
  filt8 = lpc.kautocor(blk, 8)
  assert close(list(parcor(filt8)), gold["parcor8"])
  assert close(lsf(filt8), gold["lsf8"])
  assert parcor_stable(1 / filt8) and lsf_stable(1 / filt8)


def test_stability_and_errors():
  unstable = 1 / ZFilter([1, -2.5, 1.2])
  assert not parcor_stable(unstable) and not lsf_stable(unstable)
  assert list(parcor(ZFilter([1, .5, .25, .125]))) == [0.125, 0.19047619047619047, 0.39999999999999997]
  with pytest.raises(ParCorError):
    levinson_durbin([0, 0, 0], 2)
  with pytest.raises(ParCorError):
    list(parcor(ZFilter([1, 0.3, 1.0])))
  assert issubclass(ParCorError, ZeroDivisionError)
  with pytest.raises(ValueError, match="Unstable filter"):
    lpc.kcovar([1, 2, 4, 8, 16, 32], 2)
  with pytest.raises(ValueError, match="Filter has feedback"):
    list(parcor(1 / (1 - 0.3 * z ** -1)))
  ld = levinson_durbin([4, 2], 3)                               # order past the data: zero lags appended
  assert len(ld.numerator) <= 4


@pytest.mark.gpu
def test_analysis_synthesis_on_gpu(vectors):
  """The whitening FIR (order 12, generic kernel) and its all-pole inverse applied on the GPU,
  against the reference's outputs for the same float32 input."""
  import torch
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  blk = vectors["lpc_blk"]
  filt = lpc.kautocor(blk.astype(np.float64).tolist(), 12)
  resid = list(filt(blk.tolist()))
  assert rel_err(resid, vectors["lpc_residual_y"]) <= 1e-5
  synth = list((1 / filt)(vectors["lpc_residual_y"].astype(np.float32).tolist()))
  assert rel_err(synth, vectors["lpc_synth_y"]) <= 1e-5
  assert rel_err(synth, blk) <= 1e-5                            # analysis then synthesis restores the block
