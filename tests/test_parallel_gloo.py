"""Host-side multi-GPU logic on CPU: world_size-2 gloo process group, the oracle standing
in for the device compute (injected), both sharding modes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from audiolazy_b200.parallel import split_range


def test_split_range_is_a_balanced_partition():
  for n in (0, 1, 7, 64, 65, 4096):
    for world in (1, 2, 3, 8):
      parts = [split_range(n, world, r) for r in range(world)]
      assert parts[0][0] == 0 and parts[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
      sizes = [b - a for a, b in parts]
      assert max(sizes) - min(sizes) <= 1
  with pytest.raises(ValueError):
    split_range(4, 2, 2)


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _worker(rank, world, port, out):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    import oracle
    import audiolazy_b200 as ab
    from audiolazy_b200.parallel import ShardedBank
    bank = ab.gammatone_bank(freqs=ab.erb_space(n=6), strategy="slaney")
    full_sections = bank.sections()

    def compute(local_bank, x):   # the oracle stands in for the CUDA path (no GPU in this test)
      return torch.from_numpy(oracle.bank_apply_f32(x.numpy(), local_bank.sections()))

    S, T = 5, 300
    x_all = torch.from_numpy(np.random.default_rng(0).uniform(-1, 1, (S, T)).astype(np.float32))
    want = torch.from_numpy(oracle.bank_apply_f32(x_all.numpy(), full_sections))

    # channels mode: broadcast the input block, outputs stay sharded, optional gather
    sb = ShardedBank(bank, mode="channels", compute=compute)
    x = x_all.clone() if rank == 0 else torch.zeros_like(x_all)
    sb.broadcast_input(x, src=0)
    assert torch.equal(x, x_all)
    y_local = sb.apply(x)
    assert y_local.shape == (S, sb.c_hi - sb.c_lo, T)
    assert torch.equal(y_local, want[:, sb.c_lo:sb.c_hi])
    assert torch.equal(sb.gather_output(y_local), want)
    # in-place gather: ONE all_gather_into_tensor into the rank-major [world][S][C/world][T] buffer
    assert sb.even_channels
    buf = torch.empty((world, S, 6 // world, T), dtype=torch.float32)
    sb.gather_output_into(y_local.contiguous(), buf)
    assert torch.equal(buf.permute(1, 0, 2, 3).reshape(S, 6, T), want)
    # two independent batches through the sharded bank: state=None starts from zero EVERY call (as FilterBank.apply)
    assert torch.equal(sb.apply(x), y_local)

    # streams mode: scatter rows from rank 0, no collective in apply
    sb = ShardedBank(bank, mode="streams", compute=compute)
    lo, hi = sb.stream_range(S)
    x_loc = sb.scatter_input(x_all if rank == 0 else None, S, T, src=0)
    assert torch.equal(x_loc, x_all[lo:hi])
    y_local = sb.apply(x_loc)
    assert torch.equal(y_local, want[lo:hi])
    assert torch.equal(sb.gather_output(y_local), want)
    # equal shares: one scatter straight into the destination rows
    x6 = torch.cat([x_all, x_all[:1]])                     # 6 streams, 3 per rank
    dst = torch.empty((3, T), dtype=torch.float32)
    sb.scatter_input_into(x6 if rank == 0 else None, dst, src=0)
    assert torch.equal(dst, x6[3 * rank: 3 * rank + 3])
    y6 = sb.apply(dst)
    assert torch.equal(sb.gather_output(y6), torch.from_numpy(oracle.bank_apply_f32(x6.numpy(), full_sections)))
    out.put((rank, "ok"))
  except Exception as exc:  # pragma: no cover
    out.put((rank, repr(exc)))
  finally:
    dist.destroy_process_group()


def test_world_size_2_gloo():
  ctx = mp.get_context("spawn")
  out = ctx.Queue()
  port = _free_port()
  procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
  for p in procs:
    p.start()
  results = [out.get(timeout=180) for _ in procs]
  for p in procs:
    p.join(timeout=60)
  assert sorted(results) == [(0, "ok"), (1, "ok")], results
