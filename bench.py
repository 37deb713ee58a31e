#!/usr/bin/env python
"""Benchmark of the hot path: input samples/s through the 64-channel gammatone ERB bank.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--strategy slaney]

* A STEP is one pass of the bank over one resident batch of synthetic float32 streams:
  per GPU 4096 streams x 16384 samples (BASELINE.json config 4: "64-channel gammatone ERB
  bank x 4096 independent input streams, 1 GPU HBM-bound"); with N GPUs every rank owns its
  own 4096 streams (stream sharding, weak scaling, no data-path collective; 8 ranks move the
  same 5.4e8 input samples per step as config 5's 65536 x 8192).
* ``value`` = input stream-samples/s of the whole job, device-timed with CUDA events on the
  launching stream, exactly K steps between barrier + synchronize, max over ranks.
* ``e2e`` = the same metric through the C-ABI host entry (``alz_apply_f32_host``) with PINNED
  HOST buffers: host->device copy of x and device->host copy of every output row inside the
  timed region, every step.
* ``roofline`` = algorithmic HBM bytes (260 B per input sample: 4 read + 64 x 4 written) over
  the measured launch duration, against the measured copy peak of MEASURED_PEAKS.json.
* ``cpu_baseline`` (rank 0, N = 1) and ``--impl reference``: the CPU restatement of the
  reference's evaluator (oracle/, kind "port": the reference itself is pure Python and lives
  only in the build container) on all host threads, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "samples/sec through 64-ch gammatone bank"
UNIT = "input-samples/s"
S_PER_GPU, T, C, RATE = 4096, 16384, 64, 48000
BYTES_PER_IN_SAMPLE = 4 + 4 * C


def parse():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=20)
  ap.add_argument("--warmup", type=int, default=5)
  ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
  ap.add_argument("--strategy", default="slaney", choices=["slaney", "klapuri", "sampled"])
  ap.add_argument("--streams", type=int, default=S_PER_GPU, help="streams per GPU")
  ap.add_argument("--samples", type=int, default=T, help="samples per stream per step")
  ap.add_argument("--no-e2e", action="store_true")
  ap.add_argument("--no-cpu", action="store_true")
  return ap.parse_args()


def peaks():
  try:
    with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
      return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
  except Exception:
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
  """DRAM bytes per launch of the headline kernel from the committed ncu capture."""
  try:
    with open(os.path.join(ROOT, "profiles", "ncu_summary.json")) as fh:
      return json.load(fh).get("dram_bytes_per_launch")
  except Exception:
    return None


class ClockSampler(object):
  """Samples SM clock / throttle reasons of one GPU during the timed region (NVML)."""

  def __init__(self, index):
    self.samples, self.reasons, self.max_mhz = [], set(), None
    self._stop = threading.Event()
    self._thread = None
    try:
      import pynvml
      pynvml.nvmlInit()
      self.nv = pynvml
      self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
      self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
    except Exception:
      self.nv = None

  def _run(self):
    nv = self.nv
    names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
             nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
             nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
             nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
             nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
    while not self._stop.is_set():
      try:
        self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
        for bit, name in names.items():
          if mask & bit:
            self.reasons.add(name)
      except Exception:
        pass
      self._stop.wait(0.004)

  def start(self):
    if self.nv is not None:
      self._thread = threading.Thread(target=self._run, daemon=True)
      self._thread.start()

  def stop(self):
    self._stop.set()
    if self._thread is not None:
      self._thread.join()
    s = sorted(self.samples)
    return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
            "samples": len(s)}


def bank_sections(strategy):
  import audiolazy_b200 as ab
  return ab.gammatone_bank(rate=RATE, strategy=strategy)


def cpu_port_throughput(bank, budget_s, threads):
  """Input samples/s of the oracle (C restatement) with `threads` host threads on a bounded
  sample of the bench workload; returns (value, sample description)."""
  import numpy as np
  import oracle
  sections = bank.sections()
  rng = np.random.default_rng(0)
  t_cpu = 2048
  probe = rng.uniform(-1, 1, (threads, t_cpu)).astype(np.float32)
  out = np.empty((threads, C, t_cpu), dtype=np.float32)
  out.fill(0)                      # touch the pages: first-touch faults are not the evaluator's cost
  t0 = time.perf_counter()
  oracle.bank_apply_f32(probe, sections, threads=threads, out=out)
  dt = time.perf_counter() - t0
  rate = threads * t_cpu / dt
  n_streams = int(max(threads, min(4096, (rate * budget_s / t_cpu) // threads * threads)))
  x = rng.uniform(-1, 1, (n_streams, t_cpu)).astype(np.float32)
  out = np.empty((n_streams, C, t_cpu), dtype=np.float32)
  out.fill(0)
  t0 = time.perf_counter()
  oracle.bank_apply_f32(x, sections, threads=threads, out=out)
  dt = time.perf_counter() - t0
  return n_streams * t_cpu / dt, "%d streams x %d samples x %d channels, %d threads, %.1f s" % (
    n_streams, t_cpu, C, threads, dt)


def python_port_throughput(bank):
  """The pure-Python statement-by-statement port (what CPython costs the reference), 1 core."""
  import numpy as np
  import oracle
  sections = bank.sections()
  x = np.random.default_rng(0).uniform(-1, 1, 1500).astype(np.float32).astype(float).tolist()
  t0 = time.perf_counter()
  for ch in sections[::8]:
    oracle.py_cascade(ch, x)
  dt = time.perf_counter() - t0
  return len(x) / (dt * 8)      # all 64 channels cost 8x the 8 sampled ones


def run_reference(args):
  """--impl reference: the CPU port on all host threads; rank 0 only."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  threads = os.cpu_count() or 1
  bank = bank_sections(args.strategy)
  values = []
  sample = ""
  for i in range(args.warmup + args.steps):
    v, sample = cpu_port_throughput(bank, 3.0 if i >= args.warmup else 0.5, threads)
    if i >= args.warmup:
      values.append(v)
  value = sum(values) / len(values)
  line = {
    "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
    "warmup": args.warmup, "ms_per_step": args.streams * args.samples / value * 1e3, "higher_is_better": True,
    "scaling": "weak", "vs_baseline": None,
    "dtype": "f64", "data": "synthetic",
    "config": {"workload": "64-ch gammatone ERB bank (%s), fs 48 kHz, CPU port of the reference evaluator on a "
                           "bounded sample of the %d x %d stream batch" % (args.strategy, args.streams, args.samples),
               "ms_per_step_note": "extrapolated from the sample to the whole batch (streams are independent: the "
                                   "cost is linear in their number)"},
    "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
    "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    "gpu_launches": 0,
  }
  print(json.dumps(line), flush=True)


def run_ours(args):
  import numpy as np
  import torch
  import torch.distributed as dist
  from audiolazy_b200 import _capi

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py needs a GPU (there is no CPU fallback)")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  distributed = world > 1
  if distributed:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's version / debug lines must not mix with the JSON line on stdout
    dist.init_process_group("nccl", device_id=dev)

  S, Tn = args.streams, args.samples
  bank = bank_sections(args.strategy)
  plan = bank.device_bank().plan
  gen = torch.Generator(device=dev)
  gen.manual_seed(1234 + rank)
  x = torch.rand((S, Tn), device=dev, generator=gen) * 2 - 1          # synthetic uniform(-1, 1) float32
  y = torch.empty((S, C, Tn), dtype=torch.float32, device=dev)
  state = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=dev)
  cur = torch.cuda.current_stream(dev).cuda_stream

  def step():
    plan.apply(x.data_ptr(), y.data_ptr(), state.data_ptr(), S, Tn, Tn, Tn, cur)

  def barrier():
    if distributed:
      dist.barrier()
    torch.cuda.synchronize(dev)

  for _ in range(max(args.warmup, 3)):
    step()
  barrier()
  sampler = ClockSampler(local)
  sampler.start()
  launches0 = _capi.launch_count()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  e0.record()
  for _ in range(args.steps):
    step()
  e1.record()
  barrier()
  launches = _capi.launch_count() - launches0
  clocks = sampler.stop()
  ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
  if distributed:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  ms_total = float(ms.item())
  ms_per_step = ms_total / args.steps
  value = world * S * Tn / (ms_per_step * 1e-3)

  # ---- end to end through the host-buffer C-ABI entry ------------------------------------
  e2e = None
  if not args.no_e2e:
    # pinned host buffers for the whole batch (17.4 GB per rank); shrink the e2e batch only if
    # the box cannot pin that much for every rank
    Se = S
    try:
      import psutil
      avail = psutil.virtual_memory().available
      per_stream = (C + 1) * Tn * 4
      Se = int(max(32, min(S, (avail * 0.4 / world) // per_stream // 32 * 32)))
    except Exception:
      pass
    xh = torch.empty((Se, Tn), dtype=torch.float32).pin_memory()
    xh.copy_(x[:Se].cpu())
    yh = torch.empty((Se, C, Tn), dtype=torch.float32).pin_memory()
    xn, yn = xh.numpy(), yh.numpy()
    state = torch.zeros(plan.state_doubles(Se), dtype=torch.float64, device=dev)
    plan.apply_host(xn, yn, state.data_ptr())                  # warm-up: allocates the staging buffers
    k_e2e = max(1, min(args.steps, 3))
    barrier()
    t0 = time.perf_counter()
    for _ in range(k_e2e):
      plan.apply_host(xn, yn, state.data_ptr())                # H2D x, kernel, D2H y: all inside, synchronous
    torch.cuda.synchronize(dev)
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if distributed:
      dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = {"value": world * Se * Tn * k_e2e / float(dt.item()), "unit": UNIT,
           "h2d_bytes_per_step": world * Se * Tn * 4, "d2h_bytes_per_step": world * Se * C * Tn * 4,
           "steps": k_e2e, "streams_per_gpu": Se,
           "note": "alz_apply_f32_host with pinned host buffers; PCIe-bound on the 256 B/sample output"}
    del xh, yh

  if rank == 0:
    peak, peak_src = peaks()
    achieved = BYTES_PER_IN_SAMPLE * S * Tn / (ms_per_step * 1e-3) / 1e9          # per GPU, GB/s
    line = {
      "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
      "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
      "dtype": "f64", "data": "synthetic",
      "config": {"workload": "64-ch gammatone ERB bank (%s, 50 Hz-20 kHz ERB-rate spaced, fs 48 kHz) x %d streams x %d "
                             "samples per GPU (BASELINE config 4 per GPU)" % (args.strategy, S, Tn),
                 "streams_per_gpu": S, "samples_per_stream": Tn, "channels": C, "sharding": "streams (no data-path collective)",
                 "io_dtype": "float32", "l2": "inputs (%.0f MB) and outputs (%.1f GB) per step exceed the 126 MB L2"
                 % (S * Tn * 4 / 1e6, S * C * Tn * 4 / 1e9),
                 "realtime_48k_streams": value / RATE},
      "clocks": clocks, "gpu_launches": int(launches),
      "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                   "traffic": ncu_traffic(), "peak_source": peak_src,
                   "kernel": "alz_biquad_tma_kernel<K=4,NB=2,MONIC=2> (two adjacent ceilings: the float64 arithmetic the parity bar needs, 3.57 ms, and the write stream in 256-byte row pieces, 3.46 ms; DESIGN.md section 3)"},
    }
    if e2e is not None:
      line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
      threads = os.cpu_count() or 1
      v, sample = cpu_port_throughput(bank, 10.0, threads)
      line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                              "python_port_1core": python_port_throughput(bank)}
    print(json.dumps(line), flush=True)
  if distributed:
    dist.barrier()
    dist.destroy_process_group()


def main():
  args = parse()
  if args.impl == "reference":
    run_reference(args)
  else:
    run_ours(args)


if __name__ == "__main__":
  main()
