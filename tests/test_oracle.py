"""Pin the CPU oracle (oracle/) against golden vectors produced by the reference itself
(tests/golden/make_golden.py). Bit-exact: the oracle restates float64 arithmetic in the
reference's operation order."""
import numpy as np
import pytest

import oracle
from conftest import signal


def test_bank_outputs_bit_exact(designs, vectors):
  x = signal(0, 8000)
  for name in ("slaney", "klapuri", "sampled"):
    bank = [designs["bank_" + name][c] for c in vectors["bank_channels"]]
    y = oracle.bank_apply(x, bank)[0]
    assert np.array_equal(y, vectors["bank_%s_y" % name]), name


def test_bank_impulse_responses(designs, vectors):
  imp = np.zeros(2000, dtype=np.float32)
  imp[0] = 1
  for name in ("slaney", "klapuri", "sampled"):
    bank = [designs["bank_" + name][c] for c in (4, 40)]
    assert np.array_equal(oracle.bank_apply(imp, bank)[0], vectors["bank_%s_impulse" % name])


def test_cfg1_cfg2(designs, vectors):
  y = oracle.bank_apply(signal(1, 48000), [[([1, 7, 2], [1, 0.5, 0.2])]])[0, 0]
  assert np.array_equal(y, vectors["cfg1_y"])
  sos = designs["cfg2_sos"]
  y = oracle.bank_apply(signal(2, 50000), [[(r[:3], r[3:]) for r in sos]])[0, 0]
  assert np.array_equal(y, vectors["cfg2_y"])


def test_memory_and_zero_seeding(designs, vectors):
  xs = signal(3, 64)
  f = ([0.5, -0.25, 2.0], [2.0, 0.5, -0.3])
  y = oracle.bank_apply(xs, [[f]], xinit=[[[0.125, 0.125]]], yinit=[[[0.75, -1.5]]])[0, 0]
  assert np.array_equal(y, vectors["seed_single_y"])
  # a memory shorter than the delay line is padded with `zero` on the LEFT (lazy_misc.py:132-160)
  y = oracle.bank_apply(xs, [[f]], xinit=[[[-0.5, -0.5]]], yinit=[[[-0.5, 0.75]]])[0, 0]
  assert np.array_equal(y, vectors["seed_short_memory_y"])
  casc = designs["seed_cascade"]
  xi = np.full((1, 3, 2), 0.25)
  yi = np.zeros((1, 3, 2))
  for k, (b, a) in enumerate(casc):
    mem = [0.3, -0.2][:len(a) - 1]
    yi[0, k, :len(a) - 1] = [0.25] * (len(a) - 1 - len(mem)) + mem
  y = oracle.bank_apply(xs, [casc], xinit=xi, yinit=yi)[0, 0]
  assert np.array_equal(y, vectors["seed_cascade_y"])


def test_gain_divisor(vectors):
  xs = signal(3, 64)
  y = oracle.bank_apply(xs, [[([1.0, 3.0], [-18.0, 9.8, 0.0, 14.3])]])[0, 0]
  assert np.array_equal(y, vectors["a0_not_one_y"])
  y = oracle.bank_apply(xs, [[([1.0, 0.0, -1.0], [-1.0, 0.5])]])[0, 0]
  assert np.array_equal(y, vectors["a0_minus_one_y"])


def test_high_order_and_combs(designs, vectors):
  xg = signal(4, 4000)
  y = oracle.bank_apply(xg, [[(designs["generic_b"], designs["generic_a"])]])[0, 0]
  assert np.array_equal(y, vectors["generic_y"])
  assert np.array_equal(oracle.bank_apply(xg, [designs["comb_fb_37_0.8"]])[0, 0], vectors["comb_fb_y"])
  assert np.array_equal(oracle.bank_apply(xg, [designs["comb_ff_100_-0.5"]])[0, 0], vectors["comb_ff_y"])


def test_parallel_sum(designs, vectors):
  xs = signal(3, 64)
  ys = oracle.bank_apply(xs, designs["parallel"])[0]
  acc = ys[0]
  for row in ys[1:]:
    acc = acc + row
  assert np.array_equal(acc, vectors["parallel_y"])


def test_python_port_matches_golden_and_c(designs, vectors):
  xs = signal(3, 64).astype(np.float64).tolist()
  f = ([0.5, -0.25, 2.0], [2.0, 0.5, -0.3])
  assert np.array_equal(oracle.py_section(*f, xs, memory=[0.75, -1.5], zero=0.125), vectors["seed_single_y"])
  assert np.array_equal(oracle.py_section(*f, xs, memory=[0.75], zero=-0.5), vectors["seed_short_memory_y"])
  assert np.array_equal(oracle.py_cascade(designs["seed_cascade"], xs, memory=[0.3, -0.2], zero=0.25),
                        vectors["seed_cascade_y"])
  rng = np.random.default_rng(11)
  for _ in range(20):
    nb, na = rng.integers(1, 6), rng.integers(1, 5)
    b = rng.uniform(-1, 1, nb).round(2).tolist()
    a = [float(rng.choice([1.0, -1.0, 2.0, 0.5]))] + (rng.uniform(-0.4, 0.4, na - 1)).round(2).tolist()
    x = signal(int(rng.integers(0, 1000)), 200)
    yc = oracle.bank_apply(x, [[(b, a)]])[0, 0]
    yp = np.array(oracle.py_section(b, a, x.astype(np.float64).tolist()))
    assert np.array_equal(yc, yp), (b, a)


def test_threads_and_f32_variant(designs):
  x = np.stack([signal(50 + i, 500) for i in range(6)])
  bank = designs["bank_slaney"][:8]
  y1 = oracle.bank_apply(x, bank)
  y4 = oracle.bank_apply(x, bank, threads=4)
  assert np.array_equal(y1, y4)
  yf = oracle.bank_apply_f32(x, bank, threads=3)
  assert np.array_equal(yf, y1.astype(np.float32))


def test_zero_gain_is_rejected_by_python_port():
  with pytest.raises(ZeroDivisionError):
    oracle.py_section([1.0], [0.0, 1.0], [1.0, 2.0])
