"""First end-to-end check on a B200: parity of the CUDA path vs the oracle/golden vectors
and a first timing of the headline configuration (scratch tool, superseded by tests/ and bench.py)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import oracle
from audiolazy_b200 import _capi

d = json.load(open("tests/golden/designs.json")); v = np.load("tests/golden/vectors.npz")
sig = lambda seed, n: np.random.default_rng(seed).uniform(-1, 1, n).astype(np.float32)
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
cur = lambda: torch.cuda.current_stream().cuda_stream

def gpu_apply(plan, x, xinit=None, yinit=None, splits=None):
  x = np.atleast_2d(x); S, T = x.shape; C = plan.n_channels
  xd = torch.from_numpy(x).to(dev)
  y = torch.empty((S, C, T), dtype=torch.float32, device=dev)
  st = torch.empty(max(1, plan.state_doubles(S)), dtype=torch.float64, device=dev)
  plan.state_init(st.data_ptr(), S, xinit, yinit, cur())
  splits = splits or [T]
  t0 = 0
  for n in splits:
    plan.apply(xd.data_ptr() + 4 * t0, y.data_ptr() + 4 * t0, st.data_ptr(), S, n, T, T, cur())
    t0 += n
  torch.cuda.synchronize()
  return y.cpu().numpy()

def relerr(y, ref):
  ref = np.asarray(ref, dtype=np.float64)
  return float(np.max(np.abs(y.astype(np.float64) - ref), axis=-1).max() / np.abs(ref).max()), \
         float((np.max(np.abs(y.astype(np.float64) - ref), axis=-1) / np.max(np.abs(ref), axis=-1)).max())

x = np.stack([sig(0, 8000), sig(7, 8000)])
for name in ["slaney", "klapuri", "sampled"]:
  bank = d["bank_" + name]
  plan = _capi.Plan(bank)
  print(name, "kind", plan.kind, "K", plan.n_sections, "NB", plan.num_taps, "monic", plan.monic, "ops", plan.fp64_ops)
  y = gpu_apply(plan, x)
  yo = oracle.bank_apply(x, bank)
  print("  vs oracle (all 64 ch, 2 streams): worst per-row rel err", relerr(y, yo)[1])
  g = v["bank_%s_y" % name]
  print("  vs golden:", relerr(y[0][v["bank_channels"]], g)[1])
  y2 = gpu_apply(plan, x, splits=[1, 1, 30, 33, 935, 7000])
  print("  block split bit-exact:", np.array_equal(y, y2))

plan = _capi.Plan([[([1, 7, 2], [1, 0.5, 0.2])]])
y = gpu_apply(plan, sig(1, 48000)); print("cfg1 rel err", relerr(y[0, 0], v["cfg1_y"]))
sos = d["cfg2_sos"]; bank2 = [[(r[:3], r[3:]) for r in sos]]
plan = _capi.Plan(bank2); print("cfg2 monic", plan.monic, "K", plan.n_sections, "NB", plan.num_taps)
y = gpu_apply(plan, sig(2, 50000)); print("cfg2 rel err vs golden", relerr(y[0, 0], v["cfg2_y"]))
x2 = sig(2, 1000000)
t = time.time(); yo = oracle.bank_apply(x2, bank2); tor = time.time() - t
xd = torch.from_numpy(x2[None]).to(dev); yd = torch.empty((1, 1, 1000000), dtype=torch.float32, device=dev)
st = torch.zeros(plan.state_doubles(1), dtype=torch.float64, device=dev)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
plan.apply(xd.data_ptr(), yd.data_ptr(), st.data_ptr(), 1, 1000000, 1000000, 1000000, cur()); torch.cuda.synchronize()
st.zero_(); e0.record(); plan.apply(xd.data_ptr(), yd.data_ptr(), st.data_ptr(), 1, 1000000, 1000000, 1000000, cur()); e1.record(); torch.cuda.synchronize()
print("cfg2 1e6: rel err vs oracle", relerr(yd.cpu().numpy()[0, 0], yo[0, 0]), "gpu ms", e0.elapsed_time(e1), "oracle s", tor)

xs = sig(3, 64)
plan = _capi.Plan([[([0.5, -0.25, 2.0], [2.0, 0.5, -0.3])]])
y = gpu_apply(plan, xs, xinit=[[[0.125, 0.125]]], yinit=[[[0.75, -1.5]]]); print("seed single", relerr(y[0, 0], v["seed_single_y"]))
sc = d["seed_cascade"]; plan = _capi.Plan([sc]); print("seed cascade K", plan.n_sections, "NB", plan.num_taps, "monic", plan.monic)
xi = np.full((1, plan.n_sections, 2), 0.25); yi = np.zeros((1, plan.n_sections, 2))
for k, (b, a) in enumerate(sc):
  mem = [0.3, -0.2][:len(a) - 1]; mem = [0.25] * (len(a) - 1 - len(mem)) + mem
  yi[0, k, :len(mem)] = mem
y = gpu_apply(plan, xs, xinit=xi, yinit=yi); print("seed cascade", relerr(y[0, 0], v["seed_cascade_y"]))
y2 = gpu_apply(plan, xs, xinit=xi, yinit=yi, splits=[1, 1, 1, 61]); print("  split bit-exact", np.array_equal(y, y2))
xg = sig(4, 4000)
for nm, bank, key in [("generic", [[(d["generic_b"], d["generic_a"])]], "generic_y"), ("comb fb", [d["comb_fb_37_0.8"]], "comb_fb_y"), ("comb ff", [d["comb_ff_100_-0.5"]], "comb_ff_y")]:
  plan = _capi.Plan(bank); y = gpu_apply(plan, xg); print(nm, "kind", plan.kind, relerr(y[0, 0], v[key]))
  y2 = gpu_apply(plan, xg, splits=[1, 5, 100, 3894]); print("  split bit-exact", np.array_equal(y, y2))
# ragged channel counts: C = 1, 3, 5, 48 with many streams
for C, S in [(1, 100), (3, 37), (5, 64), (48, 3), (64, 5)]:
  bank = d["bank_slaney"][:C]; plan = _capi.Plan(bank); xx = np.stack([sig(100 + i, 1000) for i in range(S)])
  y = gpu_apply(plan, xx); yo = oracle.bank_apply(xx, bank); print("C", C, "S", S, "rel err", relerr(y, yo)[1])
# unaligned pointers / strides (scalar paths)
plan = _capi.Plan(d["bank_slaney"][:64]); S, T = 3, 1001
xx = np.stack([sig(200 + i, T) for i in range(S)])
xd = torch.zeros(S * 1003 + 1, dtype=torch.float32, device=dev); yd = torch.zeros(S * 64 * 1005 + 1, dtype=torch.float32, device=dev)
xd[1:].view(S, 1003)[:, :T] = torch.from_numpy(xx).to(dev)
st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=dev)
plan.apply(xd.data_ptr() + 4, yd.data_ptr() + 4, st.data_ptr(), S, T, 1003, 1005, cur()); torch.cuda.synchronize()
y = yd[1:].view(S * 64, 1005)[:, :T].cpu().numpy().reshape(S, 64, T); print("unaligned rel err", relerr(y, oracle.bank_apply(xx, d["bank_slaney"]))[1])
# host path
yh = plan.apply_host(xx); print("host path equal to device path:", np.array_equal(yh, y))

# ---------------------------------------------------------------- timing
def time_bank(name, S, T, iters=5):
  plan = _capi.Plan(d["bank_" + name]); C = 64
  g = torch.Generator(device=dev); g.manual_seed(0)
  xd = (torch.rand((S, T), device=dev, generator=g) * 2 - 1)
  yd = torch.empty((S, C, T), dtype=torch.float32, device=dev)
  st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=dev)
  for _ in range(3): plan.apply(xd.data_ptr(), yd.data_ptr(), st.data_ptr(), S, T, T, T, cur())
  torch.cuda.synchronize(); ts = []
  for _ in range(iters):
    e0.record(); plan.apply(xd.data_ptr(), yd.data_ptr(), st.data_ptr(), S, T, T, T, cur()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
  ms = float(np.median(ts)); insamp = S * T / (ms * 1e-3)
  print("%s S=%d T=%d: %.3f ms  %.3f G in-samples/s  %.1f GB/s algorithmic (%.1f%% of 6584.5)" % (name, S, T, ms, insamp / 1e9, insamp * 260 / 1e9, insamp * 260 / 6584.5e9 * 100), ts)
  del yd
time_bank("slaney", 4096, 16384)
time_bank("slaney", 2072, 16384)
time_bank("slaney", 1036, 16384)
time_bank("klapuri", 4096, 16384)
time_bank("sampled", 256, 4096, iters=2)
# e2e host path
plan = _capi.Plan(d["bank_slaney"]); S, T = 512, 16384
xh = torch.empty((S, T), dtype=torch.float32).pin_memory(); xh.uniform_(-1, 1)
yh = torch.empty((S, 64, T), dtype=torch.float32).pin_memory()
xn, yn = xh.numpy(), yh.numpy()
plan.apply_host(xn, yn); t = time.time(); plan.apply_host(xn, yn); dt = time.time() - t
print("host e2e S=%d T=%d: %.1f ms, %.3f G in-samples/s, D2H %.1f GB/s" % (S, T, dt * 1e3, S * T / dt / 1e9, S * 64 * T * 4 / dt / 1e9))
print("launches", _capi.launch_count())
