"""Timing of the few-stream configs (BASELINE configs 2 and 3) with and without the time-parallel path."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiolazy_b200 import _capi
d = json.load(open("tests/golden/designs.json"))
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
cur = lambda: torch.cuda.current_stream().cuda_stream
def run(name, bank, T=1000000, S=1):
  plan = _capi.Plan(bank); C = plan.n_channels
  x = torch.rand((S, T), device=dev) * 2 - 1
  y = torch.empty((S, C, T), dtype=torch.float32, device=dev)
  st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=dev)
  e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
  for mode in ("1", "0"):
    os.environ["ALZ_NO_TIME_PARALLEL"] = mode
    ts = []
    for _ in range(4):
      st.zero_(); e0.record(); plan.apply(x.data_ptr(), y.data_ptr(), st.data_ptr(), S, T, T, T, cur()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("%-34s %s: %.3f ms  %.1f M input-samples/s" % (name, "sequential   " if mode == "1" else "time-parallel", min(ts), S * T / min(ts) / 1e3))
run("cfg2 butterworth-8, 1 stream", [[(r[:3], r[3:]) for r in d["cfg2_sos"]]])
run("cfg3 slaney bank, 1 stream", d["bank_slaney"])
run("cfg3 sampled bank, 1 stream", d["bank_sampled"])
run("slaney bank, 16 streams x 1e6", d["bank_slaney"], S=16)
run("slaney bank, 256 x 65536", d["bank_slaney"], T=65536, S=256)
run("slaney bank, 64 x 262144", d["bank_slaney"], T=262144, S=64)
run("cfg1-like biquad, 1 x 48000", [[([1, 7, 2], [1, 0.5, 0.2])]], T=48000)
