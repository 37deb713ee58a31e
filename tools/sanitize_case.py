"""One launch that takes the time-segmented + paired-tile path of the TMA engine, compared with the
plain path; small enough to run under compute-sanitizer:
  compute-sanitizer --tool memcheck python tools/sanitize_case.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from audiolazy_b200 import _capi

S, T, C = 2048 + 5, 2048 + 77, 64          # 65 groups x 64 channels = 4160 warps > 3552 slots; ragged edges
d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "designs.json")))
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
plan = _capi.Plan(d["bank_slaney"])
x = torch.rand((S, T + 3), device=dev)[:, :T] * 2 - 1          # row stride T + 3: unaligned rows -> cp.async engine too
xa = x.contiguous()
cur = torch.cuda.current_stream().cuda_stream
outs = []
for env in ({}, {"ALZ_NO_SEGMENT": "1", "ALZ_TMA_PAIRED": "0"}, {"ALZ_NO_TMA": "1"}):
  os.environ.update(env)
  y = torch.full((S, C, T + 3), float("nan"), device=dev)
  st = torch.zeros(plan.state_doubles(S), dtype=torch.float64, device=dev)
  # aligned input (TMA needs 16-byte aligned rows: T + 3 floats is not) -> pad the row stride to a multiple of 4
  xs = torch.zeros((S, (T + 3) // 4 * 4 + 4), device=dev)
  xs[:, :T] = xa
  ys = torch.full((S, C, xs.shape[1]), float("nan"), device=dev)
  plan.apply(xs.data_ptr(), ys.data_ptr(), st.data_ptr(), S, T, xs.shape[1], xs.shape[1], cur)
  torch.cuda.synchronize()
  outs.append((ys[:, :, :T].clone(), st.clone()))
  bad = (~torch.isnan(ys[:, :, T:])).nonzero()
  if len(bad):
    print("env", env, "wrote past the end of a row at", bad[:8].tolist(), "count", len(bad), "values", ys[:, :, T:][tuple(bad[:4].T)].tolist())
    raise SystemExit(1)
  for k in env:
    del os.environ[k]
for y, st in outs[1:]:
  assert torch.equal(outs[0][0], y) and torch.equal(outs[0][1], st)
print("segmented+paired == plain == cp.async engine: ok", tuple(outs[0][0].shape))
