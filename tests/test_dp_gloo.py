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


# ---- data-parallel equivalence of the train step: every loss is a batch mean and nothing couples
# images, so the all-reduced SUM of the shard gradients times 1/world IS the global-batch gradient -------
def _dp_worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.set_num_threads(2)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import numpy as np
    import seeded_weights
    from oracle import filters as OF
    from oracle import train_step as OT
    from exposure_b200.trainer import default_cfg
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))
    P = {}
    for which in ("generator", "critic"):
      names = [str(s) for s in gold["seed_%s_varnames" % which]]
      shapes = [tuple(int(v) for v in str(s).split(",")) for s in gold["seed_%s_varshapes" % which]]
      P.update({k: seeded_weights.make(k, s, seed=7) for k, s in zip(names, shapes)})     # identical replicas
    Pg = {k: v for k, v in P.items() if k.startswith("generator/")}
    Pv = {k: v for k, v in P.items() if k.startswith("rl_value/")}
    Pc = {k: v for k, v in P.items() if k.startswith("critic/")}
    cfg = default_cfg()
    B = 4
    g = torch.Generator().manual_seed(3)
    img = OF.synth_images(B, 64, 64, seed=5, stress=False).double() * 3
    states = torch.zeros(B, 11, dtype=torch.float64)
    states[:, 2] = torch.tensor([0.0, 1.0, 4.0, 2.0])
    noise = torch.rand(B, generator=g, dtype=torch.float64)
    drop_f = (torch.rand(B, 4, 4, 256, generator=g) < 0.5).double() * 2
    drop_s = (torch.rand(B, 4, 4, 256, generator=g) < 0.5).double() * 2
    b, e = shard_range(B, rank, world)
    sl = slice(b, e)
    mine = OT.generator_step(Pg, Pv, Pc, img[sl], states[sl], noise[sl], drop_f[sl], drop_s[sl], 0.3, cfg)
    names = sorted(mine["grads_g"])
    flat = torch.cat([mine["grads_g"][k].reshape(-1) for k in names] + [mine["grads_v"][k].reshape(-1) for k in sorted(mine["grads_v"])])
    dist.all_reduce(flat)                                   # ONE all-reduce of the flat gradient buffer ...
    flat = flat * (1.0 / world)                             # ... and the mean factor the fused Adam applies
    loss = torch.stack([mine["g_loss"], mine["v_loss"]])
    dist.all_reduce(loss)
    loss = loss / world
    if rank == 0:
      full = OT.generator_step(Pg, Pv, Pc, img, states, noise, drop_f, drop_s, 0.3, cfg)
      ref = torch.cat([full["grads_g"][k].reshape(-1) for k in names] + [full["grads_v"][k].reshape(-1) for k in sorted(full["grads_v"])])
      rel = float((flat - ref).abs().max() / ref.abs().max())
      lrel = float((loss - torch.stack([full["g_loss"], full["v_loss"]])).abs().max())
      q.put((rel, lrel, int(flat.numel())))
  finally:
    dist.destroy_process_group()


def test_sharded_generator_step_equals_global_batch_gradient():
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29100 + os.getpid() % 500
  procs = [ctx.Process(target=_dp_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  rel, lrel, n = q.get(timeout=300)
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  assert n == 6123680 + 1221857            # theta_g + theta_v (one fused 29.4 MB buffer)
  assert rel < 1e-12 and lrel < 1e-12, (rel, lrel)
