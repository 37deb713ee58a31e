"""Multi-GPU check on real NCCL (run under torchrun): channel-sharded bank = NCCL broadcast of the
input block + local channels + (separately timed) all_gather of the outputs; stream-sharded bank =
scatter of the input rows.  Prints parity against the single-GPU result and the measured collective
bandwidths (the gather is the NVLink-bound step SURVEY.md 8e keeps off the throughput path)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import audiolazy_b200 as ab
from audiolazy_b200.parallel import ShardedBank

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bank = ab.gammatone_bank(strategy="slaney")
S, T = 1024, 16384
g = torch.Generator(device=dev); g.manual_seed(7)
x_all = torch.rand((S, T), device=dev, generator=g) * 2 - 1          # same seed on every rank
full = bank.apply(x_all) if rank == 0 else None

def timed(fn, n=5):
  fn(); torch.cuda.synchronize(); dist.barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(n): fn()
  e1.record(); torch.cuda.synchronize()
  ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
  return float(ms.item())

# ---- channels mode -------------------------------------------------------------------------
sb = ShardedBank(bank, mode="channels")
x = x_all.clone() if rank == 0 else torch.zeros_like(x_all)
sb.broadcast_input(x, src=0)
assert torch.equal(x, x_all)
y_local = sb.apply(x)
y_gathered = sb.gather_output(y_local)
if rank == 0:
  assert torch.equal(y_gathered, full), "channel-sharded result differs"
ms_b = timed(lambda: sb.broadcast_input(x, src=0))
ms_k = timed(lambda: sb.apply(x))
ms_g = timed(lambda: sb.gather_output(y_local), n=2)
# ---- streams mode --------------------------------------------------------------------------
ss = ShardedBank(bank, mode="streams")
lo, hi = ss.stream_range(S)
x_loc = ss.scatter_input(x_all if rank == 0 else None, S, T, src=0, device=dev, dtype=torch.float32)
assert torch.equal(x_loc, x_all[lo:hi])
y_loc = ss.apply(x_loc)
ok = torch.tensor([1], device=dev)
if rank == 0:
  ok[0] = int(torch.equal(y_loc, full[lo:hi]))
dist.broadcast(ok, src=0)
if rank == 0:
  in_bytes, out_bytes = S * T * 4, S * (64 // world) * T * 4
  print("nccl_check world=%d: parity ok=%d; channels mode: broadcast %.3f ms (%.1f GB/s), kernel %.3f ms (%.2f G in-samples/s per rank-step), "
        "all_gather %.3f ms (%.1f GB/s algorithmic per rank)" % (world, int(ok.item()), ms_b, in_bytes / ms_b / 1e6, ms_k, S * T / ms_k / 1e6,
        ms_g, out_bytes * (world - 1) / ms_g / 1e6))
dist.barrier()
dist.destroy_process_group()
