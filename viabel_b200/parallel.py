"""Multi-GPU plumbing for the sharded hot path: one process per GPU,
torch.distributed over NCCL (gloo in the CPU tests of this host logic).

* ELBO gradient: observations are sharded by rows; every rank draws identical base variates
  (same Philox seed / offset), so the only exchange is a sum of the per-rank sweep outputs
  [ll(S), gmu(d), ge(d)] -- a <= 10 KB, latency-bound all-reduce.
* PSIS: draws are sharded; ranks exchange their candidate lists (all-gather) and a handful of
  scalars; the smoothed weights stay sharded.
"""
import torch
import torch.distributed as dist

__all__ = ['is_distributed', 'world', 'shard_rows', 'allreduce_sum_', 'allgather_ragged']


def is_distributed(group=None):
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def world(group=None):
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_rows(n, rank, world_size):
    """Contiguous, balanced [lo, hi) slice of n rows for `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError('rank out of range')
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t, group=None):
    """In-place sum over ranks (no-op for a single process).  Returns t."""
    if is_distributed(group):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def allgather_ragged(t, group=None):
    """Concatenate 1-D tensors of different lengths from all ranks (rank order)."""
    if not is_distributed(group):
        return t
    ws = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    cap = max(sizes) if sizes else 0
    pad = torch.zeros(cap, dtype=t.dtype, device=t.device)
    pad[:t.numel()] = t
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)])
