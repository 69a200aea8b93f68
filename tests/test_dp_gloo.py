"""CPU / gloo (world_size 2) coverage of the N>1 host logic: batch sharding, flat gradient
all-reduce + mean factor, identical initial weights, distinct per-rank random streams and replay
shards.  (The CUDA kernels are not involved: ParamStore and ReplayMemory are plain torch tensors.)"""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from exposure_b200.dp import allreduce_grads, rank_seed, shard_range


def test_shard_range_partitions_exactly():
  for n in (0, 1, 7, 64, 65, 257):
    for world in (1, 2, 3, 8):
      spans = [shard_range(n, r, world) for r in range(world)]
      assert spans[0][0] == 0 and spans[-1][1] == n
      assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
      sizes = [e - b for b, e in spans]
      assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from exposure_b200.nets import CriticNet, ParamStore
    from exposure_b200.replay import ReplayMemory, SyntheticProvider
    from exposure_b200.trainer import default_cfg
    dev = torch.device("cpu")
    store = ParamStore(dev)
    CriticNet(store, "critic", n_states=0)
    store.finalize(seed=5)                       # same seed on every rank -> identical replicas
    w = [torch.zeros_like(store.flat) for _ in range(world)]
    dist.all_gather(w, store.flat)
    same_init = all(torch.equal(w[0], t) for t in w)
    store.grad.fill_(float(rank + 1))
    scale = allreduce_grads(store)
    ok_sum = bool(torch.all(store.grad == float(sum(range(1, world + 1)))))
    cfg = default_cfg()
    cfg.batch_size = 8
    cfg.replay_memory_size = 16
    mem = ReplayMemory(cfg, SyntheticProvider(dev, "raw", rank_seed(1, rank, 0), size=64),
                       SyntheticProvider(dev, "real", rank_seed(1, rank, 1), size=64), dev, seed=rank_seed(1, rank, 2))
    img, st, slots = mem.get_next_fake_batch(8)
    digest = torch.tensor([float(img.sum())])
    d = [torch.zeros(1) for _ in range(world)]
    dist.all_gather(d, digest)
    distinct = len({float(t) for t in d}) == world
    q.put((rank, same_init, ok_sum, scale, distinct))
  finally:
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_and_replicas():
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + os.getpid() % 500
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=120) for _ in range(world)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, same_init, ok_sum, scale, distinct in res:
    assert same_init, "weight replicas differ across ranks"
    assert ok_sum, "all-reduce did not sum the flat gradient buffer"
    assert scale == 0.5
    assert distinct, "ranks drew identical replay batches"
