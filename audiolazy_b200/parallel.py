"""Multi-GPU sharding of a filterbank: one process per GPU, ``torch.distributed`` (NCCL
over NVLink on the B200 box, gloo in CPU tests) for the plumbing.

Every (stream, channel) pair is an independent recurrence, so the path shards with no
data-path collective in steady state:

* ``mode="streams"`` -- rank ``r`` owns streams ``[lo, hi)`` and all channels. Inputs that
  already live on their rank need no communication at all (the benchmark's weak-scaling
  configuration); :meth:`ShardedBank.scatter_input` distributes a batch that starts on
  one rank.
* ``mode="channels"`` -- rank ``r`` owns channels ``[lo, hi)`` of every stream; the input
  block is broadcast (4 B per input sample per receiving GPU), outputs stay sharded.
  :meth:`ShardedBank.gather_output` (all_gather of ``S * C_local * T`` floats per rank) is
  provided for consumers that need everything in one place; it is NVLink-bound and is
  deliberately not part of :meth:`ShardedBank.apply`.

The compute callable is injectable so that the host-side logic is testable on CPU with
the gloo backend.
"""
from __future__ import annotations

__all__ = ["split_range", "ShardedBank"]


def split_range(n, world, rank):
  """Contiguous, balanced partition of ``range(n)``: ``(lo, hi)`` of ``rank``."""
  if world < 1 or not 0 <= rank < world:
    raise ValueError("bad world/rank")
  base, extra = divmod(n, world)
  lo = rank * base + min(rank, extra)
  return lo, lo + base + (1 if rank < extra else 0)


class ShardedBank(object):
  """A :class:`~audiolazy_b200.bank.FilterBank` sharded over the ranks of a process group."""

  def __init__(self, bank, mode="streams", group=None, compute=None):
    import torch.distributed as dist
    if mode not in ("streams", "channels"):
      raise ValueError("mode must be 'streams' or 'channels'")
    self.mode = mode
    self.group = group
    self.dist = dist
    self.world = dist.get_world_size(group) if dist.is_initialized() else 1
    self.rank = dist.get_rank(group) if dist.is_initialized() else 0
    self.bank = bank
    self.n_channels = len(bank)
    if mode == "channels":
      self.c_lo, self.c_hi = split_range(self.n_channels, self.world, self.rank)
      from .bank import FilterBank
      self.local = FilterBank(list(bank)[self.c_lo:self.c_hi])
    else:
      self.c_lo, self.c_hi = 0, self.n_channels
      self.local = bank
    self._compute = compute
    self._state = None

  # -- partitioning ----------------------------------------------------------------------
  def stream_range(self, n_streams):
    return split_range(n_streams, self.world, self.rank) if self.mode == "streams" else (0, n_streams)

  # -- communication ---------------------------------------------------------------------
  def broadcast_input(self, x, src=0):
    """channels mode: every rank needs the whole input block (in place broadcast)."""
    if self.world > 1:
      self.dist.broadcast(x, src=src, group=self.group)
    return x

  def scatter_input(self, x_full, n_streams, n_samples, src=0, device=None, dtype=None):
    """streams mode: rank ``src`` holds ``x_full[S, T]``; returns this rank's rows."""
    import torch
    lo, hi = self.stream_range(n_streams)
    if self.world == 1:
      return x_full[lo:hi]
    ref = x_full if x_full is not None else None
    device = device or (ref.device if ref is not None else "cpu")
    dtype = dtype or (ref.dtype if ref is not None else torch.float32)
    out = torch.empty((hi - lo, n_samples), dtype=dtype, device=device)
    # equal-sized chunks are required by scatter: pad the ragged tail
    per = -(-n_streams // self.world)
    buf = torch.zeros((per, n_samples), dtype=dtype, device=device)
    chunks = None
    if self.rank == src:
      chunks = []
      for r in range(self.world):
        a, b = split_range(n_streams, self.world, r)
        piece = torch.zeros((per, n_samples), dtype=dtype, device=device)
        piece[: b - a] = x_full[a:b]
        chunks.append(piece)
    self.dist.scatter(buf, chunks, src=src, group=self.group)
    out.copy_(buf[: hi - lo])
    return out

  def gather_output(self, y_local):
    """Everything everywhere: ``y[S, C, T]`` (channels mode) or ``y[S_total, C, T]``
    (streams mode) on every rank. NVLink-bound; time it separately."""
    import torch
    if self.world == 1:
      return y_local
    if self.mode == "channels":
      sizes = [split_range(self.n_channels, self.world, r) for r in range(self.world)]
      per = max(b - a for a, b in sizes)
      S, _, T = y_local.shape
      pad = torch.zeros((S, per, T), dtype=y_local.dtype, device=y_local.device)
      pad[:, : y_local.shape[1]] = y_local
      parts = [torch.empty_like(pad) for _ in range(self.world)]
      self.dist.all_gather(parts, pad, group=self.group)
      return torch.cat([p[:, : b - a] for p, (a, b) in zip(parts, sizes)], dim=1)
    counts = [torch.zeros(1, dtype=torch.int64, device=y_local.device) for _ in range(self.world)]
    self.dist.all_gather(counts, torch.tensor([y_local.shape[0]], dtype=torch.int64, device=y_local.device),
                         group=self.group)
    counts = [int(c.item()) for c in counts]
    per = max(counts)
    _, C, T = y_local.shape
    pad = torch.zeros((per, C, T), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    parts = [torch.empty_like(pad) for _ in range(self.world)]
    self.dist.all_gather(parts, pad, group=self.group)
    return torch.cat([p[:n] for p, n in zip(parts, counts)], dim=0)

  # -- compute ---------------------------------------------------------------------------
  def apply(self, x_local, state=None, out=None):
    """Filter this rank's shard: ``x_local[S_local, T]`` -> ``y[S_local, C_local, T]``.
    No collective is issued here."""
    if self._compute is not None:
      return self._compute(self.local, x_local)
    if state is None:
      if self._state is None or self._state.n_streams != x_local.shape[0]:
        self._state = self.local.new_state(x_local.shape[0])
      state = self._state
    return self.local.apply(x_local, state=state, out=out)
