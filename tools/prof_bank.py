"""Run the bank kernel a few times (for ncu / quick timing): python tools/prof_bank.py slaney 4096 16384 3"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiolazy_b200 import _capi
name, S, T, iters = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
C = int(sys.argv[5]) if len(sys.argv) > 5 else 64
d = json.load(open("tests/golden/designs.json"))
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
if name.startswith("first"):      # "first1".."first4": only the first n sections of the slaney cascades (engine ceiling probes)
  n = int(name[5:])
  plan = _capi.Plan([ch[:n] for ch in d["bank_slaney"][:C]])
else:
  plan = _capi.Plan(d["bank_" + name][:C])
xd = torch.rand((S, T), device=dev) * 2 - 1
yd = torch.empty((S, C, T), dtype=torch.float32, device=dev)
st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=dev)
cur = torch.cuda.current_stream().cuda_stream
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
ts = []
major = os.environ.get("ALZ_PROF_LAYOUT", "stream")       # "channel": y[C][S][T] through alz_apply_f32_ex
for i in range(iters):
  e0.record()
  if major == "channel":
    plan.apply_ex(xd.data_ptr(), yd.data_ptr(), st.data_ptr(), S, T, T, S * T, T, cur)
  else:
    plan.apply(xd.data_ptr(), yd.data_ptr(), st.data_ptr(), S, T, T, T, cur)
  e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
ms = min(ts)
print("[%s-major] " % major + "%s C=%d S=%d T=%d best %.3f ms  %.3f G in-samples/s  %.1f GB/s" % (name, C, S, T, ms, S * T / ms / 1e6, S * T * (4 + 4 * C) / ms / 1e6), ["%.3f" % t for t in ts])
