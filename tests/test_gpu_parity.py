"""Parity of the CUDA path (through the C ABI) with the oracle and with the golden vectors
the reference produced. Tolerance: the north star's <= 1e-5 relative (to each output row's
peak) for float32 I/O; measured errors are ~6e-8 (float32 rounding of the float64 result)."""
import numpy as np
import pytest

import oracle
from conftest import rel_err, signal

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def gpu():
  import torch
  if not torch.cuda.is_available():
    pytest.skip("no CUDA device")
  torch.cuda.set_device(0)
  from audiolazy_b200 import _capi
  assert _capi.device_count() >= 1

  class G:
    pass

  g = G()
  g.torch = torch
  g.capi = _capi
  g.dev = torch.device("cuda:0")

  def run(plan, x, xinit=None, yinit=None, splits=None):
    x = np.atleast_2d(np.asarray(x, dtype=np.float32))
    S, T = x.shape
    C = plan.n_channels
    xd = torch.from_numpy(x).to(g.dev)
    y = torch.full((S, C, max(T, 1)), float("nan"), dtype=torch.float32, device=g.dev)[:, :, :T].contiguous()
    st = torch.empty(max(1, plan.state_doubles(S)), dtype=torch.float64, device=g.dev)
    cur = torch.cuda.current_stream().cuda_stream
    plan.state_init(st.data_ptr(), S, xinit, yinit, cur)
    t0 = 0
    for n in (splits or [T]):
      plan.apply(xd.data_ptr() + 4 * t0, y.data_ptr() + 4 * t0, st.data_ptr(), S, n, max(T, 1), max(T, 1), cur)
      t0 += n
    torch.cuda.synchronize()
    return y.cpu().numpy()

  g.run = run
  return g


@pytest.mark.parametrize("name", ["slaney", "klapuri", "sampled"])
def test_bank_vs_golden_and_oracle(gpu, designs, vectors, name):
  bank = designs["bank_" + name]
  plan = gpu.capi.Plan(bank)
  assert plan.kind == gpu.capi.KIND_BIQUAD            # sampled: head-FIR variant (8-tap first section)
  assert plan.num_taps == {"slaney": 2, "klapuri": 3, "sampled": 8}[name]
  x = np.stack([signal(0, 8000), signal(7, 8000), signal(8, 8000)])
  y = gpu.run(plan, x)
  assert rel_err(y[0][vectors["bank_channels"]], vectors["bank_%s_y" % name]) <= TOL      # the reference's own output
  assert rel_err(y, oracle.bank_apply(x, bank)) <= TOL                                    # all 64 channels, 3 streams
  imp = np.zeros(2000, dtype=np.float32)
  imp[0] = 1
  yi = gpu.run(plan, imp)
  assert rel_err(yi[0][[4, 40]], vectors["bank_%s_impulse" % name]) <= TOL


def test_cfg1_cfg2(gpu, designs, vectors):
  plan = gpu.capi.Plan([[([1, 7, 2], [1, 0.5, 0.2])]])
  assert rel_err(gpu.run(plan, signal(1, 48000))[0, 0], vectors["cfg1_y"]) <= TOL
  bank2 = [[(r[:3], r[3:]) for r in designs["cfg2_sos"]]]
  plan = gpu.capi.Plan(bank2)
  assert plan.monic and plan.n_sections == 4
  assert rel_err(gpu.run(plan, signal(2, 50000))[0, 0], vectors["cfg2_y"]) <= TOL
  x = signal(2, 1000000)                         # BASELINE cfg 2 at full size, against the oracle
  assert rel_err(gpu.run(plan, x)[0, 0], oracle.bank_apply(x, bank2)[0, 0]) <= TOL


def test_memory_zero_seeding(gpu, designs, vectors):
  xs = signal(3, 64)
  f = ([0.5, -0.25, 2.0], [2.0, 0.5, -0.3])
  plan = gpu.capi.Plan([[f]])
  y = gpu.run(plan, xs, xinit=[[[0.125, 0.125]]], yinit=[[[0.75, -1.5]]])
  assert rel_err(y[0, 0], vectors["seed_single_y"]) <= TOL
  y = gpu.run(plan, xs, xinit=[[[-0.5, -0.5]]], yinit=[[[-0.5, 0.75]]])
  assert rel_err(y[0, 0], vectors["seed_short_memory_y"]) <= TOL
  casc = designs["seed_cascade"]
  plan = gpu.capi.Plan([casc])
  K = plan.n_sections
  xi = np.zeros((1, K, 2))
  yi = np.zeros((1, K, 2))
  for k, (b, a) in enumerate(casc):
    xi[0, k, :len(b) - 1] = 0.25
    mem = [0.3, -0.2][:len(a) - 1]
    yi[0, k, :len(a) - 1] = [0.25] * (len(a) - 1 - len(mem)) + mem
  y = gpu.run(plan, xs, xinit=xi, yinit=yi)
  assert rel_err(y[0, 0], vectors["seed_cascade_y"]) <= TOL
  for splits in ([1, 63], [1, 1, 62], [2, 62], [33, 31]):
    assert np.array_equal(gpu.run(plan, xs, xinit=xi, yinit=yi, splits=splits), y)


def test_gain_divisor_and_generic_kernel(gpu, designs, vectors):
  xs = signal(3, 64)
  y = gpu.run(gpu.capi.Plan([[([1.0, 3.0], [-18.0, 9.8, 0.0, 14.3])]]), xs)
  assert rel_err(y[0, 0], vectors["a0_not_one_y"]) <= TOL
  y = gpu.run(gpu.capi.Plan([[([1.0, 0.0, -1.0], [-1.0, 0.5])]]), xs)
  assert rel_err(y[0, 0], vectors["a0_minus_one_y"]) <= TOL
  xg = signal(4, 4000)
  for bank, key in [([[(designs["generic_b"], designs["generic_a"])]], "generic_y"),
                    ([designs["comb_fb_37_0.8"]], "comb_fb_y"), ([designs["comb_ff_100_-0.5"]], "comb_ff_y")]:
    plan = gpu.capi.Plan(bank)
    assert plan.kind == gpu.capi.KIND_GENERIC
    y = gpu.run(plan, xg)
    assert rel_err(y[0, 0], vectors[key]) <= TOL
    assert np.array_equal(gpu.run(plan, xg, splits=[1, 5, 100, 3894]), y)     # ring buffers are block-exact


@pytest.mark.parametrize("C,S,T", [(1, 1, 1), (1, 1, 2), (1, 3, 3), (1, 100, 1000), (3, 37, 257), (5, 64, 31), (5, 33, 32),
                                   (48, 3, 33), (64, 5, 999), (64, 32, 64), (7, 65, 100)])
def test_ragged_shapes(gpu, designs, C, S, T):
  bank = designs["bank_slaney"][:C]
  plan = gpu.capi.Plan(bank)
  x = np.stack([signal(100 + i, T) for i in range(S)])
  assert rel_err(gpu.run(plan, x), oracle.bank_apply(x, bank)) <= TOL


def test_block_split_is_bit_exact(gpu, designs):
  x = np.stack([signal(0, 8000), signal(7, 8000)])
  for name in ("slaney", "klapuri", "sampled"):
    plan = gpu.capi.Plan(designs["bank_" + name])
    y = gpu.run(plan, x)
    assert np.array_equal(gpu.run(plan, x, splits=[1, 1, 30, 33, 935, 7000]), y)
    assert np.array_equal(gpu.run(plan, x, splits=[4000, 4000]), y)


def test_unaligned_pointers_and_strides(gpu, designs):
  torch = gpu.torch
  bank = designs["bank_slaney"]
  plan = gpu.capi.Plan(bank)
  S, T = 3, 1001
  xx = np.stack([signal(200 + i, T) for i in range(S)])
  xd = torch.zeros(S * 1003 + 1, dtype=torch.float32, device=gpu.dev)
  yd = torch.zeros(S * 64 * 1005 + 1, dtype=torch.float32, device=gpu.dev)
  xd[1:].view(S, 1003)[:, :T] = torch.from_numpy(xx).to(gpu.dev)
  st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=gpu.dev)
  plan.apply(xd.data_ptr() + 4, yd.data_ptr() + 4, st.data_ptr(), S, T, 1003, 1005, torch.cuda.current_stream().cuda_stream)
  torch.cuda.synchronize()
  y = yd[1:].view(S * 64, 1005)[:, :T].cpu().numpy().reshape(S, 64, T)
  assert rel_err(y, oracle.bank_apply(xx, bank)) <= TOL
  assert np.array_equal(y, gpu.run(plan, xx))          # scalar and vector paths agree bit for bit


def test_host_path_and_state_carry(gpu, designs):
  torch = gpu.torch
  bank = designs["bank_klapuri"][:16]
  plan = gpu.capi.Plan(bank)
  x = np.stack([signal(300 + i, 5000) for i in range(9)])
  want = gpu.run(plan, x)
  assert np.array_equal(plan.apply_host(x), want)
  st = torch.zeros(plan.state_doubles(9), dtype=torch.float64, device=gpu.dev)
  a = plan.apply_host(np.ascontiguousarray(x[:, :1234]), state_ptr=st.data_ptr())
  b = plan.apply_host(np.ascontiguousarray(x[:, 1234:]), state_ptr=st.data_ptr())
  assert np.array_equal(np.concatenate([a, b], axis=2), want)


def test_parallel_channel_sum(gpu, designs, vectors):
  torch = gpu.torch
  plan = gpu.capi.Plan(designs["parallel"])
  xs = signal(3, 64)
  y = torch.from_numpy(gpu.run(plan, xs)).to(gpu.dev)
  out = torch.empty((1, 64), dtype=torch.float32, device=gpu.dev)
  gpu.capi.sum_channels(y.data_ptr(), out.data_ptr(), 1, 3, 64, 64, 64, torch.cuda.current_stream().cuda_stream)
  torch.cuda.synchronize()
  assert rel_err(out.cpu().numpy()[0], vectors["parallel_y"]) <= TOL


def test_error_codes(gpu):
  with pytest.raises(ZeroDivisionError):
    gpu.capi.Plan([[([1.0], [0.0, 1.0])]])
  with pytest.raises(ValueError):
    gpu.capi.Plan([[([float("nan")], [1.0])]])
  plan = gpu.capi.Plan([[([1.0], [1.0])], []])        # second channel: empty cascade = identity
  x = signal(5, 100)
  y = gpu.run(plan, x)
  assert np.array_equal(y[0, 0], x) and np.array_equal(y[0, 1], x)


def test_full_size_properties(gpu, designs, monkeypatch):
  """BASELINE cfg 4 (64 ch x 4096 streams x 16384 samples), where the oracle cannot go:
  stream independence, block-split exactness, time-segmented == unsegmented launch, and spot
  rows against the oracle."""
  torch = gpu.torch
  bank = designs["bank_slaney"]
  plan = gpu.capi.Plan(bank)
  S, T, C = 4096, 16384, 64
  g = torch.Generator(device=gpu.dev)
  g.manual_seed(0)
  x = torch.rand((S, T), device=gpu.dev, generator=g) * 2 - 1
  x[S // 2:] = x[:S // 2]                     # second half duplicates the first
  cur = torch.cuda.current_stream().cuda_stream
  y = torch.empty((S, C, T), dtype=torch.float32, device=gpu.dev)
  st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=gpu.dev)
  plan.apply(x.data_ptr(), y.data_ptr(), st.data_ptr(), S, T, T, T, cur)
  torch.cuda.synchronize()
  assert torch.equal(y[:S // 2], y[S // 2:])   # same input, same output, whatever the lane / warp / SM
  rows = [0, 1, 31, 32, 2047, 777]
  want = oracle.bank_apply(x[rows].cpu().numpy(), bank)
  assert rel_err(y[rows].cpu().numpy(), want) <= TOL
  y2 = torch.empty_like(y)
  st.zero_()
  half = T // 2 + 32 * 3 + 5
  plan.apply(x.data_ptr(), y2.data_ptr(), st.data_ptr(), S, half, T, T, cur)
  plan.apply(x.data_ptr() + 4 * half, y2.data_ptr() + 4 * half, st.data_ptr(), S, T - half, T, T, cur)
  torch.cuda.synchronize()
  assert torch.equal(y, y2)
  assert bool(torch.isfinite(y[::97]).all())
  # the launches above were time-segmented (8192 warps = 2.3 waves); one plain launch must agree bit for bit
  monkeypatch.setenv("ALZ_NO_SEGMENT", "1")
  y2.fill_(float("nan"))
  st2 = torch.zeros_like(st)
  plan.apply(x.data_ptr(), y2.data_ptr(), st2.data_ptr(), S, T, T, T, cur)
  torch.cuda.synchronize()
  monkeypatch.delenv("ALZ_NO_SEGMENT")
  assert torch.equal(y, y2)
  assert torch.equal(st, st2)


def test_time_parallel_path(gpu, designs, monkeypatch):
  """Few long streams (BASELINE configs 2 and 3) take the chunked zero-state / scan / replay
  path; it must agree with the sequential path and the oracle, honour seeds and carry state."""
  x = np.stack([signal(70 + i, 200000 + 77) for i in range(3)])
  for bank, xi, yi in [(designs["bank_slaney"][:64], None, None),
                       ([[(r[:3], r[3:]) for r in designs["cfg2_sos"]]], np.full((1, 4, 2), 0.25), np.full((1, 4, 2), -0.5)),
                       (designs["bank_sampled"][:5], None, None)]:
    plan = gpu.capi.Plan(bank)
    fast = gpu.run(plan, x, xinit=xi, yinit=yi)
    monkeypatch.setenv("ALZ_NO_TIME_PARALLEL", "1")
    slow = gpu.run(plan, x, xinit=xi, yinit=yi)
    monkeypatch.delenv("ALZ_NO_TIME_PARALLEL")
    assert rel_err(fast, slow) <= 3e-6          # rounding of the chunk states only (sampled: ill-conditioned head FIR)
    want = oracle.bank_apply(x[:1, :150000], bank, xinit=xi, yinit=yi)
    assert rel_err(fast[:1, :, :150000], want) <= TOL
    # state carry: a long block (chunked) followed by short ones (sequential) == one shot
    split = gpu.run(plan, x, xinit=xi, yinit=yi, splits=[131072 + 5, 1000, 200077 - 131077 - 1000])
    assert rel_err(split, fast) <= 3e-6
  # a dozen streams: one batch
  x = np.stack([signal(300 + i, 70000) for i in range(12)])
  plan = gpu.capi.Plan(designs["bank_slaney"][:16])
  fast = gpu.run(plan, x)
  monkeypatch.setenv("ALZ_NO_TIME_PARALLEL", "1")
  slow = gpu.run(plan, x)
  monkeypatch.delenv("ALZ_NO_TIME_PARALLEL")
  assert rel_err(fast, slow) <= 3e-6


@pytest.mark.parametrize("S,T,C", [(33, 40000, 64), (100, 16384 + 37, 16), (256, 65536, 8), (1000, 20000, 1), (5, 300001, 64)])
def test_time_parallel_batches_all_streams(gpu, designs, monkeypatch, S, T, C):
  """33 ... ~1000 streams that would leave most of the GPU idle are evaluated time-parallel in ONE batch (virtual
  streams = (stream, chunk) pairs, 3-D / 4-D tensor maps; the cp.async engine when rows are not 16-byte aligned):
  against the sequential evaluation and, on a few rows, against the oracle; state carried into a following block."""
  bank = designs["bank_slaney"][:C]
  plan = gpu.capi.Plan(bank)
  rng = np.random.default_rng(S)
  x = rng.uniform(-1, 1, (S, T)).astype(np.float32)
  fast = gpu.run(plan, x, splits=[T - 1000, 1000])       # long block: time-parallel; short block: sequential, from its state
  monkeypatch.setenv("ALZ_NO_TIME_PARALLEL", "1")
  slow = gpu.run(plan, x)
  monkeypatch.delenv("ALZ_NO_TIME_PARALLEL")
  assert rel_err(fast, slow) <= 3e-6
  rows = [0, S // 2, S - 1]
  assert rel_err(fast[rows][:, :, :30000], oracle.bank_apply(x[rows][:, :30000], bank)) <= TOL
  seq = gpu.capi.Plan(bank, sequential=True)             # ALZ_PLAN_SEQUENTIAL: any blocking, same bits
  assert np.array_equal(gpu.run(seq, x), slow)
  assert np.array_equal(gpu.run(seq, x, splits=[T - 1000, 1000]), slow)


def test_host_path_time_segments(gpu, designs):
  """A single long stream whose output exceeds one staging chunk is cut into time segments
  (state carried on the device): same result as the device path."""
  bank = designs["bank_slaney"]
  plan = gpu.capi.Plan(bank)
  x = signal(90, 600000)[None, :]                     # 64 ch x 600000 x 4 B = 154 MB > 128 MiB
  want = gpu.run(plan, x)
  got = plan.apply_host(x)
  assert rel_err(got, want) <= 3e-6                    # time-parallel chunking differs between the two splits
  assert rel_err(got[:, :, :20000], oracle.bank_apply(x[:, :20000], bank)) <= TOL


def test_gain_modes(gpu, designs, monkeypatch):
  """MONIC mode 2 (float32 input-side gain, default) and mode 1 (float64 output gain, ALZ_EXACT_GAIN=1)."""
  bank = designs["bank_slaney"]
  x = np.stack([signal(0, 8000), signal(7, 8000)])
  want = oracle.bank_apply(x, bank)
  fast = gpu.capi.Plan(bank, exact=True)              # exact: every channel on the float64 tier
  assert fast.monic and fast.fp64_ops == 12 and fast.n_fp32_channels == 0
  assert rel_err(gpu.run(fast, x), want) <= 2.5e-7
  monkeypatch.setenv("ALZ_EXACT_GAIN", "1")
  exact = gpu.capi.Plan(bank, exact=True)
  monkeypatch.delenv("ALZ_EXACT_GAIN")
  assert exact.fp64_ops == 13
  assert rel_err(gpu.run(exact, x), want) <= 6.5e-8   # = float32 rounding of the float64 result


@pytest.mark.parametrize("name", ["slaney", "klapuri", "sampled"])
def test_precision_tiers(gpu, designs, monkeypatch, name):
  """Channels whose float32 evaluation the plan-time probe measured at <= 2.5e-6 run their recurrence in
  float32 (tier 1), the others in float64 (tier 0).  The bar is 1e-5: tier-1 channels must keep a 3x margin
  on signals the probe has not seen, tier-0 channels are float32 roundings of the float64 result."""
  bank = designs["bank_" + name]
  plan = gpu.capi.Plan(bank)
  tier, probe = plan.tiers()
  assert plan.n_fp32_channels == int(tier.sum()) and 8 <= plan.n_fp32_channels <= 48
  assert abs(plan.tier_tol - 2.5e-6) < 1e-12
  assert np.all(probe[tier == 1] <= plan.tier_tol) and np.all(probe[tier == 0] > plan.tier_tol)
  assert not tier[:16].any()                           # poles next to z = 1: never float32
  x = np.stack([signal(0, 8000), signal(7, 8000), signal(8, 8000), signal(21, 8000)])
  want = oracle.bank_apply(x, bank)
  y = gpu.run(plan, x)
  err = np.max(np.max(np.abs(y - want), axis=-1) / np.max(np.abs(want), axis=-1), axis=0)      # per channel
  assert np.all(err[tier == 0] <= 2.5e-7)
  assert np.all(err[tier == 1] <= TOL / 3), err[tier == 1].max()
  exact = gpu.capi.Plan(bank, exact=True)
  assert exact.n_fp32_channels == 0 and not exact.tiers()[0].any()
  ye = gpu.run(exact, x)
  assert rel_err(ye, want) <= 2.5e-7
  assert np.array_equal(ye[:, tier == 0], y[:, tier == 0])        # the float64 channels do not depend on the tiering
  monkeypatch.setenv("ALZ_NO_FP32_TIER", "1")
  assert gpu.capi.Plan(bank).n_fp32_channels == 0
  monkeypatch.delenv("ALZ_NO_FP32_TIER")
  assert np.array_equal(gpu.run(plan, x, splits=[1, 1, 30, 33, 935, 7000]), y)   # block splitting stays bit-exact in both tiers


def test_structurally_zero_taps_are_skipped(gpu, designs, monkeypatch):
  """gammatone.klapuri's sections are [1 - z^-2, const] / poles twice: the numerator taps that are
  zero in every channel are not computed (10 instead of 16 float64 operations per channel-sample);
  results are identical to the kernel that multiplies by the zeros."""
  bank = designs["bank_klapuri"]
  x = np.stack([signal(0, 8000), signal(7, 8000)])
  lean = gpu.capi.Plan(bank)
  assert lean.fp64_ops == 10
  monkeypatch.setenv("ALZ_NO_ZMASK", "1")
  full = gpu.capi.Plan(bank)
  monkeypatch.delenv("ALZ_NO_ZMASK")
  assert full.fp64_ops == 16
  y = gpu.run(lean, x, splits=[3000, 5000])
  assert np.array_equal(y, gpu.run(full, x))
  assert rel_err(y, oracle.bank_apply(x, bank)) <= TOL


def test_row_padding_is_never_written(gpu, designs):
  """Rows wider than n_samples (16-byte aligned strides, so the TMA engine runs): nothing outside
  [0, n_samples) of an output row may be written, whatever n_samples % 4 is (the TMA clips a box at
  16-byte granularity, so ragged ends are stored by the lanes)."""
  torch = gpu.torch
  bank = designs["bank_slaney"][:5]
  plan = gpu.capi.Plan(bank)
  cur = torch.cuda.current_stream().cuda_stream
  for S, T in [(37, 2125), (3, 77), (33, 33), (2, 5), (40, 64 + 2), (5, 96)]:
    x = np.stack([signal(500 + i, T) for i in range(S)])
    want = gpu.run(plan, x)
    stride = (T + 3) // 4 * 4 + 8
    xs = torch.zeros((S, stride), device=gpu.dev)
    xs[:, :T] = torch.from_numpy(x).to(gpu.dev)
    ys = torch.full((S, len(bank), stride), float("nan"), device=gpu.dev)
    st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=gpu.dev)
    plan.apply(xs.data_ptr(), ys.data_ptr(), st.data_ptr(), S, T, stride, stride, cur)
    torch.cuda.synchronize()
    assert bool(torch.isnan(ys[:, :, T:]).all()), (S, T)
    assert np.array_equal(ys[:, :, :T].cpu().numpy(), want), (S, T)


@pytest.mark.skipif(not __import__("os").environ.get("ALZ_TEST_EXPERIMENTAL"), reason="experimental engine: set ALZ_TEST_EXPERIMENTAL=1")
def test_experimental_wide_cta_engine(gpu, designs, monkeypatch):
  """ALZ_WARPS_PER_CTA=3 (alz_lane_tma_wide.cuh) must reproduce the single-warp engine bit for bit:
  plain, paired, time-segmented and ragged launches."""
  torch = gpu.torch
  cur = torch.cuda.current_stream().cuda_stream
  for name in ["bank_slaney", "bank_klapuri"]:
    plan = gpu.capi.Plan(designs[name])
    for S, T in [(40, 500), (2048 + 5, 2048 + 76), (2048 + 5, 2048 + 77), (4096, 4096)]:
      x = torch.rand((S, T), device=gpu.dev) * 2 - 1
      stride = (T + 3) // 4 * 4
      xs = torch.zeros((S, stride), device=gpu.dev)
      xs[:, :T] = x
      outs = []
      for wide in ("1", "3"):
        monkeypatch.setenv("ALZ_WARPS_PER_CTA", wide)
        ys = torch.full((S, 64, stride), float("nan"), device=gpu.dev)
        st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=gpu.dev)
        plan.apply(xs.data_ptr(), ys.data_ptr(), st.data_ptr(), S, T, stride, stride, cur)
        torch.cuda.synchronize()
        outs.append((ys, st))
      monkeypatch.delenv("ALZ_WARPS_PER_CTA")
      assert torch.equal(outs[0][0][:, :, :T], outs[1][0][:, :, :T]), (name, S, T)
      assert bool(torch.isnan(outs[1][0][:, :, T:]).all()), (name, S, T)
      assert torch.equal(outs[0][1], outs[1][1]), (name, S, T)


def test_parallel_sum_in_one_kernel(gpu, designs):
  """alz_apply_sum_f32 (ParallelFilter, reference lazy_filters.py:1048-1054): float64 channel results summed left to
  right inside ONE kernel, one rounding to float32; against the oracle's float64 channel outputs summed the same way."""
  torch = gpu.torch
  for bank in (designs["bank_slaney"][:8], designs["bank_klapuri"][40:43], [[([1.0, 0.5], [1.0, -0.9])], [([-1.0, -0.5], [1.0, -0.9001])]]):
    plan = gpu.capi.Plan(bank, parallel=True)
    assert plan.kind == gpu.capi.KIND_BIQUAD and plan.n_fp32_channels == 0 and not plan.monic
    S, T = 70, 5000
    x = np.random.default_rng(len(bank)).uniform(-1, 1, (S, T)).astype(np.float32)
    ch = oracle.bank_apply(x, bank)
    want = ch[:, 0].copy()
    for c in range(1, len(bank)):
      want = want + ch[:, c]
    xd = torch.from_numpy(x).to(gpu.dev)
    out = torch.full((S, T), float("nan"), dtype=torch.float32, device=gpu.dev)
    st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=gpu.dev)
    cur = torch.cuda.current_stream().cuda_stream
    before = gpu.capi.launch_count()
    plan.apply_sum(xd.data_ptr(), out.data_ptr(), st.data_ptr(), S, T, T, T, cur)
    torch.cuda.synchronize()
    assert gpu.capi.launch_count() - before == 1
    got = out.cpu().numpy()
    assert rel_err(got, want) <= 2.5e-7          # relative to the SUMMED signal's peak, also under heavy cancellation (third bank)
    # blocks: the state buffer carries every channel between calls, bit for bit
    out2 = torch.empty_like(out)
    st.zero_()
    for t0, n in ((0, 4), (4, 36), (40, 2008), (2048, T - 2048)):        # 16-byte aligned block starts (TMA rows)
      plan.apply_sum(xd.data_ptr() + 4 * t0, out2.data_ptr() + 4 * t0, st.data_ptr(), S, n, T, T, cur)
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
  # unaligned rows are refused loudly (the caller then uses alz_apply_f32 + alz_sum_channels_f32)
  with pytest.raises(gpu.capi.NativeError):
    plan.apply_sum(xd.data_ptr() + 4, out.data_ptr(), st.data_ptr(), S, T - 1, T, T, cur)


@pytest.mark.parametrize("bits", [8, 16, 24, 32])
def test_wav_to_bank_to_chunks(gpu, designs, tmp_path, bits):
  """the formats either side of the path -- PCM wave files in (reference lazy_wav.py:31-130), float32
  chunks out (lazy_io.py:48-128) -- around the device bank.  Decoding and packing are exact; the filtering is within the bar."""
  import struct
  import wave
  import audiolazy_b200 as ab
  top = 1 << (bits - 1)
  rng = np.random.default_rng(bits)
  paths, ints = [], []
  for i in range(3):
    v = rng.integers(-top, top, 3000 + 100 * i).tolist()
    ints.append(v)
    path = str(tmp_path / ("s%d.wav" % i))
    with wave.open(path, "wb") as w:
      w.setnchannels(1); w.setsampwidth(bits // 8); w.setframerate(48000)
      if bits == 8:
        w.writeframes(bytes((q + 128) & 0xff for q in v))
      elif bits == 24:
        w.writeframes(b"".join(struct.pack("<i", q)[:3] for q in v))
      else:
        w.writeframes(struct.pack("<%d%s" % (len(v), "h" if bits == 16 else "i"), *v))
    paths.append(path)
  batch, lengths, rates = ab.wav_batch(paths)
  assert lengths == [len(v) for v in ints] and rates == [48000] * 3
  for i, v in enumerate(ints):                        # exact: float32(int / 2**(bits-1)), zero padded
    assert np.array_equal(batch[i, :len(v)], (np.asarray(v, dtype=np.float64) / top).astype(np.float32))
    assert not batch[i, len(v):].any()
    assert list(ab.WavStream(paths[i])) == [q / top for q in v]          # the lazy reader: the reference's exact float64
  bank = ab.FilterBank([ab.gammatone.slaney(f * ab.sHz(48000)[1], 0.02) for f in (300., 1200., 5000.)])
  y = bank.apply(gpu.torch.from_numpy(batch).to(gpu.dev)).cpu().numpy()
  assert rel_err(y, oracle.bank_apply(batch, bank.sections())) <= TOL
  row = y[1, 2, :lengths[1]]
  blocks = list(ab.chunks(row.tolist(), size=1024))
  assert b"".join(blocks) == np.concatenate([row, np.zeros(-len(row) % 1024, dtype=np.float32)]).astype("<f4").tobytes()


@pytest.mark.parametrize("name,mode", [("slaney", "abs"), ("klapuri", "rms"), ("sampled", "squared")])
def test_fused_envelope_consumer(gpu, name, mode):
  """alz_apply_envelope_f32[_host]: |y| / y^2 -> one-pole lowpass -> every 48th value, inside the bank kernel, against the
  unfused pipeline (bank output in float32, then the same lowpass in float64 on the host)."""
  import audiolazy_b200 as ab
  bank = ab.gammatone_bank(strategy=name)
  S, T, D = 37, 48 * 100, 48
  x = np.random.default_rng(3).uniform(-1, 1, (S, T)).astype(np.float32)
  xd = gpu.torch.from_numpy(x).to(gpu.dev)
  y = bank.apply(xd).cpu().numpy().astype(np.float64)
  g, R = bank._envelope_pole(np.pi / 512)
  r = np.abs(y) if mode == "abs" else y * y
  e = np.zeros_like(r)
  acc = np.zeros(r.shape[:2])
  for n in range(T):
    acc = g * r[:, :, n] + R * acc
    e[:, :, n] = acc
  want = (np.sqrt(e) if mode == "rms" else e)[:, :, D - 1::D]
  got = bank.envelope(xd, decim=D, mode=mode).cpu().numpy()
  assert got.shape == (S, 64, T // D)
  assert rel_err(got, want) <= TOL
  got_host = bank.envelope_host(x, decim=D, mode=mode)
  assert np.array_equal(got_host, got)


def test_cfg3_at_full_length_and_cfg5_shape(gpu, designs):
  """BASELINE config 3 at its full 10^6 samples (one stream through the 64-channel bank; the time-parallel path) against
  the oracle over the WHOLE length, and config 5's per-GPU shape (8192 streams x 8192 samples) against the oracle on a
  few rows plus the duplicated-half identity."""
  torch = gpu.torch
  bank = designs["bank_slaney"]
  plan = gpu.capi.Plan(bank)
  x = signal(33, 1000000)[None, :]
  y = gpu.run(plan, x)
  want = oracle.bank_apply(x, bank)
  assert rel_err(y, want) <= TOL
  del y, want
  S, T, C = 8192, 8192, 64
  g = torch.Generator(device=gpu.dev)
  g.manual_seed(5)
  xd = torch.rand((S, T), device=gpu.dev, generator=g) * 2 - 1
  xd[S // 2:] = xd[:S // 2]
  yd = torch.empty((S, C, T), dtype=torch.float32, device=gpu.dev)
  st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=gpu.dev)
  plan.apply(xd.data_ptr(), yd.data_ptr(), st.data_ptr(), S, T, T, T, torch.cuda.current_stream().cuda_stream)
  torch.cuda.synchronize()
  assert torch.equal(yd[:S // 2], yd[S // 2:])
  rows = [0, 31, 32, 4095, 2049, 777]
  assert rel_err(yd[rows].cpu().numpy(), oracle.bank_apply(xd[rows].cpu().numpy(), bank)) <= TOL
