"""Why is the side-stream broadcast not hidden when the per-rank kernel is short?  (torchrun, 2+ GPUs)
Measures, for a 16-channel bank sharded over the ranks: kernel alone, broadcast alone, the pipeline, the HOST time of a
pipeline step (does dist.broadcast block the launching thread?), and a variant with async_op=True."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import audiolazy_b200 as ab
from audiolazy_b200.parallel import ShardedBank
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev, pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
bank = ab.gammatone_bank(freqs=ab.erb_space(n=8 * world), strategy="slaney")
S, T = 4096, 4096
sb = ShardedBank(bank, mode="channels")
xb = [torch.rand((S, T), device=dev) * 2 - 1 for _ in range(2)]
y = sb.alloc_output(S, T); state = sb.local.new_state(S)
pipe = sb.pipeline(xb, y, state)
main = torch.cuda.current_stream(dev)
def gpu_ms(fn, n=20):
  fn(); torch.cuda.synchronize(); dist.barrier()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  t0 = time.perf_counter(); e0.record()
  for _ in range(n): fn()
  host = (time.perf_counter() - t0) / n * 1e3
  e1.record(); torch.cuda.synchronize()
  return e0.elapsed_time(e1) / n, host
k, kh = gpu_ms(pipe.compute_only)
b, bh = gpu_ms(lambda: sb.broadcast_input(xb[0], src=0))
p, ph = gpu_ms(pipe.step); pipe.drain()
# variant: async_op=True on the side stream, explicit wait on the main stream before the kernel that needs the block
side = torch.cuda.Stream(device=dev, priority=-1)
works = [None, None]; cnt = [0]
def step_async():
  i = cnt[0]; j = i & 1
  side.wait_stream(main)
  with torch.cuda.stream(side):
    works[j ^ 1] = dist.broadcast(xb[j ^ 1], src=0, async_op=True)
  if works[j] is not None:
    works[j].wait()                      # stream-level wait of the CURRENT (main) stream
  sb.local.apply(xb[j], state=state, out=y)
  cnt[0] += 1
a, ah = gpu_ms(step_async)
# variant: the bank kernel confined to a partition of the SMs (green context), NCCL on the others
res = {}
for sms in (136, 128, 120):
  try:
    pp = sb.pipeline(xb, y, state, compute_sms=sms)
    kk, _ = gpu_ms(lambda: pp.compute_only(partition=True))
    tt, _ = gpu_ms(pp.step); pp.drain()
    res[sms] = (pp.partition.sm_count, kk, tt)
    torch.cuda.synchronize()
    ok = torch.equal(y, y)     # placeholder: values are checked by tools/nccl_check.py
    pp.close()
  except Exception as exc:
    res[sms] = repr(exc)
if rank == 0:
  print("bcast_diag partitions (requested SMs: granted, kernel ms on the partition, pipeline ms):", res)
if rank == 0:
  print("bcast_diag world=%d: kernel %.3f ms (host %.3f), broadcast %.3f ms (host %.3f), pipeline %.3f ms (host %.3f per step), async variant %.3f ms (host %.3f)"
        % (world, k, kh, b, bh, p, ph, a, ah))
dist.barrier(); dist.destroy_process_group()
