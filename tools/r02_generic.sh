#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_pytest.txt; cat gpurun_out/r02_pytest.txt
python - <<'PY' 2>&1 | tee gpurun_out/r02_generic.txt
import sys, os, numpy as np, torch
sys.path.insert(0, ".")
import bench, audiolazy_b200 as ab
from audiolazy_b200 import _capi
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
S, T = 65536, 16384
rng = np.random.default_rng(5)
lpc_a = [1.0] + (rng.uniform(-1, 1, 12) * 0.5 ** np.arange(1, 13)).tolist()
cases = {"comb_fb_37": ab.comb.fb(37, .8).sections(), "comb_ff_100": ab.comb.ff(100, -.5).sections(), "lpc12_fir": [(lpc_a, [1.0])],
         "lpc12_allpole": [([1.0], lpc_a)], "order3_iir": [([1.0, .5, .2, .1], [1.0, -.3, .1, .05])],
         "biquad_ref": [([1.0, 0.5, 0.2], [1.0, -0.3, 0.1])]}
for name, secs in cases.items():
  for env in ({}, {"ALZ_NO_WINDOW": "1"}):
    os.environ.pop("ALZ_NO_WINDOW", None); os.environ.update(env)
    plan = _capi.Plan([secs])
    r = bench.device_record(torch, dev, plan, S, T, steps=3, warm=1)
    print("%-14s %-18s kind %d ops %2d: %8.3f ms %7.1f G samples/s %7.1f GB/s" % (name, "old generic" if env else "default", plan.kind, plan.fp64_ops, r["ms"], r["input_samples_per_s"] / 1e9, r["gbs"]))
PY
