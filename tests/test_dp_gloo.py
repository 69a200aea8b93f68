"""gloo (world_size 2) coverage of the N > 1 path: batch sharding, ONE flat theta_g + theta_v buffer and its all-reduce,
identical initial weights, distinct per-rank random streams and replay shards (CPU); the Trainer's own world > 1
branch on two processes sharing cuda:0 (-m gpu); data-parallel equivalence of the oracle train step (CPU)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pytest

from exposure_b200.dp import rank_seed, shard_range


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
    from exposure_b200.nets import CriticNet, ParamArena, ParamStore, PolicyNet
    from exposure_b200.replay import ReplayMemory, SyntheticProvider
    from exposure_b200.trainer import default_cfg
    dev = torch.device("cpu")
    # theta_g and theta_v share ONE flat buffer (one exchange per generator step, net.py:330-331)
    gen, val = ParamStore(dev), ParamStore(dev)
    PolicyNet(gen, n_states=11, scope="generator")
    CriticNet(val, "rl_value/critic", n_states=11)
    arena = ParamArena(dev, [gen, val], [5, 6])            # same seeds on every rank -> identical replicas
    layout_ok = (gen.flat.data_ptr() == arena.flat.data_ptr() and
                 val.flat.data_ptr() == arena.flat.data_ptr() + 4 * gen.padded() and
                 arena.numel == gen.padded() + val.padded() and arena.numel % (4 * 16) == 0 and
                 gen.count() == 6123680 and val.count() == 1221857)
    w = [torch.zeros_like(arena.flat) for _ in range(world)]
    dist.all_gather(w, arena.flat)
    same_init = all(torch.equal(w[0], t) for t in w)
    gen.g["generator/Conv/weights"].fill_(float(rank + 1))
    val.g["rl_value/critic/fully_connected_1/biases"].fill_(10.0 * (rank + 1))
    dist.all_reduce(arena.grad)                            # what Trainer._apply("gv") does on the fallback transport
    tot = float(sum(range(1, world + 1)))
    ok_sum = bool(torch.all(gen.grad[:gen.g["generator/Conv/weights"].numel()] == tot)) and \
        float(val.g["rl_value/critic/fully_connected_1/biases"]) == 10.0 * tot
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
    q.put((rank, layout_ok, same_init, ok_sum, distinct))
  finally:
    dist.destroy_process_group()


def test_two_rank_joint_buffer_allreduce_and_replicas():
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29500 + os.getpid() % 500
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=180) for _ in range(world)]
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  for rank, layout_ok, same_init, ok_sum, distinct in res:
    assert layout_ok, "theta_g / theta_v are not two slices of one padded flat buffer"
    assert same_init, "weight replicas differ across ranks"
    assert ok_sum, "all-reduce did not sum the joint gradient buffer"
    assert distinct, "ranks drew identical replay batches"


# ---- the Trainer's own world > 1 branch (graph replay -> exchange -> Adam with 1/world), two ranks --------------
def _trainer_worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.cuda.set_device(0)
  dist.init_process_group("gloo", rank=rank, world_size=world)       # both ranks share cuda:0: gloo moves CUDA tensors
  try:
    from oracle import filters as OF
    from exposure_b200.trainer import Trainer
    dev = torch.device("cuda", 0)
    t = Trainer(device=dev, seed=3)
    assert t.world == world and t._peer is None                       # not an NCCL group: dist.all_reduce transport
    B = 8
    g = torch.Generator().manual_seed(5)
    img = (OF.synth_images(B, 64, 64, seed=21, stress=False) * 3).to(dev)
    real = (OF.synth_images(B, 64, 64, seed=31, stress=False) * 6).clamp(0, 1.2).to(dev)
    states = torch.zeros(B, 11, device=dev)
    states[:, 2] = torch.tensor([0, 1, 2, 3, 4, 4, 6, 7.0], device=dev)
    noise = torch.rand(B, generator=g).to(dev)
    drop_f = ((torch.rand(B, 4, 4, 256, generator=g) < 0.5).float() * 2).to(dev)
    drop_s = ((torch.rand(B, 4, 4, 256, generator=g) < 0.5).float() * 2).to(dev)
    alpha = torch.rand(B, generator=g).to(dev)
    with torch.no_grad():
      t.cri.p["critic/fully_connected_1/weights"].mul_(40.0)          # gradient penalty active
    p0 = [t.gv.flat.clone(), t.cri.flat.clone()]
    b, e = shard_range(B, rank, world)
    sl = slice(b, e)
    out = t.generator_step(img[sl].contiguous(), states[sl].contiguous(), noise[sl].contiguous(), drop_f[sl].contiguous(),
                           drop_s[sl].contiguous(), 0.3, lr_g=1e-5, apply=True)
    fake = torch.cat([o.clone() for o in _gather(out["fake_output"], world)])    # the global batch of generator outputs
    t.critic_step(real[sl].contiguous(), fake[sl].contiguous(), alpha[sl].contiguous(), lr_c=1e-5, apply=True)
    torch.cuda.synchronize()
    mean_gv, mean_c = t.gv.grad / world, t.cri.grad / world           # the exchanged SUM times the factor Adam applies
    after = [t.gv.flat.clone(), t.cri.flat.clone()]
    reps = [_gather(a, world) for a in after]
    replicas_equal = all(torch.equal(r[0], x) for r in reps for x in r)
    res = None
    if rank == 0:
      # the same two steps on the GLOBAL batch by a single-process trainer (world forced to 1)
      ref = Trainer(device=dev, seed=3)
      ref.world, ref._peer = 1, None
      with torch.no_grad():
        ref.cri.p["critic/fully_connected_1/weights"].mul_(40.0)
      ref.generator_step(img, states, noise, drop_f, drop_s, 0.3, lr_g=1e-5, apply=True)
      ref.critic_step(real, fake, alpha, lr_c=1e-5, apply=True)
      torch.cuda.synchronize()
      rel = lambda a, r: float((a - r).abs().max() / (r.abs().max() + 1e-30))
      # parameters: a first Adam step moves every variable by ~lr * sign(g), so compare in units of that step and count
      # outliers (a gradient that is pure rounding noise may flip sign between the two summation orders)
      def off(a, r, p):
        step = float((r - p).abs().max())
        return float(((a - r).abs() > 0.05 * step).float().mean()), step
      res = (rel(mean_gv, ref.gv.grad), rel(mean_c, ref.cri.grad), off(after[0], ref.gv.flat, p0[0]), off(after[1], ref.cri.flat, p0[1]))
    q.put((rank, replicas_equal, res))
    dist.barrier()
  finally:
    dist.destroy_process_group()


def _gather(t, world):
  out = [torch.zeros_like(t) for _ in range(world)]
  dist.all_gather(out, t.contiguous())
  return out


@pytest.mark.gpu
def test_trainer_two_ranks_equal_the_global_batch_step(built_lib):
  """drives Trainer._apply itself with world_size 2 (two processes on cuda:0, gloo): shard
  gradients exchanged over the joint theta_g + theta_v buffer and over theta_c, mean folded into Adam; replicas
  stay identical and equal the single-process step on the global batch (gradients to 1e-4 of each buffer's max,
  parameters to a fraction of one Adam step)."""
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29300 + os.getpid() % 500
  procs = [ctx.Process(target=_trainer_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  res = [q.get(timeout=300) for _ in range(world)]
  for p in procs:
    p.join(timeout=120)
    assert p.exitcode == 0
  for rank, replicas_equal, r in res:
    assert replicas_equal, "replicas diverged after one data-parallel step"
    if r is not None:
      g_gv, g_c, (frac_gv, step_gv), (frac_c, step_c) = r
      assert g_gv < 1e-4 and g_c < 1e-4, (g_gv, g_c)
      assert step_gv > 0 and step_c > 0
      assert frac_gv < 2e-3 and frac_c < 2e-3, r


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
