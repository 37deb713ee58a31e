"""Timing of the few-stream configs (BASELINE configs 2 and 3) with and without the time-parallel path."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiolazy_b200 import _capi
d = json.load(open("tests/golden/designs.json"))
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
cur = lambda: torch.cuda.current_stream().cuda_stream
def run(name, bank, T=1000000):
  plan = _capi.Plan(bank); C = plan.n_channels
  x = torch.rand((1, T), device=dev) * 2 - 1
  y = torch.empty((1, C, T), dtype=torch.float32, device=dev)
  st = torch.zeros(plan.state_doubles(1), dtype=torch.float64, device=dev)
  e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
  for mode in ("1", "0"):
    os.environ["ALZ_NO_TIME_PARALLEL"] = mode
    ts = []
    for _ in range(4):
      st.zero_(); e0.record(); plan.apply(x.data_ptr(), y.data_ptr(), st.data_ptr(), 1, T, T, T, cur()); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print("%-28s %s: %.3f ms  %.1f M input-samples/s" % (name, "sequential   " if mode == "1" else "time-parallel", min(ts), T / min(ts) / 1e3))
run("cfg2 butterworth-8, 1 stream", [[(r[:3], r[3:]) for r in d["cfg2_sos"]]])
run("cfg3 slaney bank, 1 stream", d["bank_slaney"])
run("cfg3 sampled bank, 1 stream", d["bank_sampled"])
