"""Host-side data-parallel helpers (one process per GPU, torch.distributed).

The hot path shards by image (every filter parameter and every CNN activation is per-image,
SURVEY 8e): rank r owns a contiguous slice of the batch, holds a full weight replica and its own
replay-pool shard and RNG streams.  The only collective is ONE all-reduce (sum) of each
optimizer's flat gradient buffer per step; the mean (1/world) is folded into the fused Adam."""
import torch
import torch.distributed as dist


def world_info():
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def shard_range(n, rank, world):
  """Contiguous, balanced [begin, end) slice of n items for `rank` (sizes differ by <= 1)."""
  base, rem = divmod(n, world)
  begin = rank * base + min(rank, rem)
  return begin, begin + base + (1 if rank < rem else 0)


def allreduce_grads(store):
  """Sum the flat gradient buffer of one optimizer across ranks (no-op for world == 1).
  Returns the factor the optimizer must apply to turn the sum into the global-batch mean."""
  _, world = world_info()
  if world > 1:
    dist.all_reduce(store.grad)
  return 1.0 / world


def rank_seed(base, rank, stream=0):
  """Distinct, reproducible RNG seeds per (rank, stream): dropout, z, alpha, data order."""
  return int(base) * 1000003 + int(rank) * 101 + int(stream)
