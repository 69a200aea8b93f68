"""Data-parallel check on N real GPUs (torchrun --nproc-per-node N tools/dp_check.py): the peer-memory
all-reduce + Adam kernel (exp_dp_allreduce_adam) against a single-process step on the global batch, replica
equality, and the same steps replayed from CUDA graphs.  Prints one line per check; exit code 1 on failure."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exposure_b200.dp import shard_range   # noqa: E402
from oracle import filters as OF           # noqa: E402


def gather(t, world):
  out = [torch.zeros_like(t) for _ in range(world)]
  dist.all_gather(out, t.contiguous())
  return out


def main():
  rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  dist.init_process_group("nccl", device_id=dev)
  from exposure_b200.trainer import Trainer
  t = Trainer(device=dev, seed=3)
  ok = True
  say = lambda *a: print("[dp_check rank %d]" % rank, *a, flush=True)
  if rank == 0:
    say("transport:", "peer-memory all-reduce + Adam kernel" if t._peer is not None else "dist.all_reduce fallback")
  B = 8 * world
  g = torch.Generator().manual_seed(5)
  img = (OF.synth_images(B, 64, 64, seed=21, stress=False) * 3).to(dev)
  real = (OF.synth_images(B, 64, 64, seed=31, stress=False) * 6).clamp(0, 1.2).to(dev)
  states = torch.zeros(B, 11, device=dev)
  states[:, 2] = (torch.arange(B) % 5).float().to(dev)
  noise = torch.rand(B, generator=g).to(dev)
  drop_f = ((torch.rand(B, 4, 4, 256, generator=g) < 0.5).float() * 2).to(dev)
  drop_s = ((torch.rand(B, 4, 4, 256, generator=g) < 0.5).float() * 2).to(dev)
  alpha = torch.rand(B, generator=g).to(dev)
  with torch.no_grad():
    t.cri.p["critic/fully_connected_1/weights"].mul_(40.0)
  p0 = [t.gv.flat.clone(), t.cri.flat.clone()]
  b, e = shard_range(B, rank, world)
  sl = slice(b, e)
  c = lambda x: x[sl].contiguous()
  out = t.generator_step(c(img), c(states), c(noise), c(drop_f), c(drop_s), 0.3, lr_g=1e-5, apply=True)
  fake = torch.cat([o.clone() for o in gather(out["fake_output"], world)])
  t.critic_step(c(real), c(fake), c(alpha), lr_c=1e-5, apply=True)
  torch.cuda.synchronize()
  after = [t.gv.flat.clone(), t.cri.flat.clone()]
  for name, a in zip(("theta_g+theta_v", "theta_c"), after):
    reps = gather(a, world)
    same = all(torch.equal(reps[0], x) for x in reps)
    ok &= same
    if rank == 0:
      say("replicas bit-identical after one step (%s):" % name, same)
  if rank == 0:
    os.environ["EXPOSURE_DP_TRANSPORT"] = "none"        # the reference trainer lives on rank 0 alone: no collective set-up
    ref = Trainer(device=dev, seed=3)
    os.environ["EXPOSURE_DP_TRANSPORT"] = "peer"
    ref.world, ref._peer = 1, None
    with torch.no_grad():
      ref.cri.p["critic/fully_connected_1/weights"].mul_(40.0)
    ref.generator_step(img, states, noise, drop_f, drop_s, 0.3, lr_g=1e-5, apply=True)
    ref.critic_step(real, fake, alpha, lr_c=1e-5, apply=True)
    torch.cuda.synchronize()
    for name, a, r, p in (("theta_g+theta_v", after[0], ref.gv.flat, p0[0]), ("theta_c", after[1], ref.cri.flat, p0[1])):
      step = float((r - p).abs().max())
      frac = float(((a - r).abs() > 0.05 * step).float().mean())
      say("%s vs single-process global-batch step: max Adam step %.3e, fraction of parameters off by > 5%% of it: %.2e"
          % (name, step, frac))
      ok &= frac < 2e-3 and step > 0
  dist.barrier()
  # the same schedule from CUDA graphs (the exchange kernel is a node of the graph), a few iterations
  from exposure_b200.replay import ReplayMemory, SyntheticProvider
  from exposure_b200.trainer import default_cfg
  cfg = default_cfg()
  cfg.batch_size = 16
  cfg.replay_memory_size = 32
  t2 = Trainer(cfg, dev, seed=0)
  mem = ReplayMemory(cfg, SyntheticProvider(dev, "raw", 100 + rank), SyntheticProvider(dev, "real", 200 + rank), dev, seed=rank)
  t2.attach_memory(mem, torch.Generator(device=dev).manual_seed(300 + rank))
  t2.train_iteration(0, giters=12, citers=1)
  t2.enable_graphs(16)
  if rank == 0:
    say("graphs captured, optimizer inside the graph:", t2._graph_apply)
  for it in range(1, 6):
    o = t2.train_iteration(it, giters=1, citers=5)
  torch.cuda.synchronize()
  for name, a in (("theta_g+theta_v", t2.gv.flat), ("theta_c", t2.cri.flat)):
    reps = gather(a, world)
    same = all(torch.equal(reps[0], x) for x in reps)
    ok &= same
    if rank == 0:
      say("replicas bit-identical after 5 graph-replayed iterations (%s):" % name, same,
          "finite:", bool(torch.isfinite(a).all()))
    ok &= bool(torch.isfinite(a).all())
  # the exchange kernels alone (no profiler: CUDA events, 30 back-to-back launches, max over ranks): what one optimizer
  # step adds at this world size -- compare with exp_adam alone (the single-GPU optimizer step)
  if t2._peer is not None:
    from exposure_b200 import nn_ops as K
    res = {}
    for which, buf in (("gv", t2.gv), ("c", t2.cri)):
      for _ in range(3):
        t2._apply(which)
      dist.barrier()
      torch.cuda.synchronize()
      e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
      e0.record()
      for _ in range(30):
        t2._apply(which)
      e1.record()
      torch.cuda.synchronize()
      tt = torch.tensor([e0.elapsed_time(e1) / 30], device=dev, dtype=torch.float64)
      dist.all_reduce(tt, op=dist.ReduceOp.MAX)
      e0.record()
      for _ in range(30):
        K.adam(buf.flat, buf.grad, buf.m, buf.v, t2._hyper["c"], 0.5, 0.9, 1e-8, 1.0)
      e1.record()
      torch.cuda.synchronize()
      res[which] = (float(tt), e0.elapsed_time(e1) / 30, buf.grad.numel() * 4 / 1e6)
    if rank == 0:
      for which, (dp_ms, adam_ms, mb) in res.items():
        say("exchange+Adam kernel, %s buffer (%.1f MB), world %d: %.1f us per launch (exp_adam alone: %.1f us) -> %.0f GB/s of "
            "NVLink traffic per GPU" % (which, mb, world, dp_ms * 1e3, adam_ms * 1e3, 2 * (world - 1) / world * mb / dp_ms))
      per_iter = res["gv"][0] + 5 * res["c"][0] - (res["gv"][1] + 5 * res["c"][1])
      say("per iteration (1 generator + 5 critic steps) the exchange adds %.0f us over the single-GPU optimizer steps" % (per_iter * 1e3))
  flag = torch.tensor([1 if ok else 0], device=dev)
  dist.all_reduce(flag, op=dist.ReduceOp.MIN)
  if rank == 0:
    say("DP CHECK", "OK" if int(flag) else "FAILED")
  dist.barrier()
  dist.destroy_process_group()
  sys.exit(0 if int(flag) else 1)


if __name__ == "__main__":
  main()
