"""CPU: sharding of the 4K chain workload (BASELINE configs[4]): the plan, and -- on two gloo ranks -- that
row-sharded parameter-gradient partials add up to the whole image's gradient (linearity of finalize + regressor)."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from exposure_b200.chain import shard_plan
from oracle import filters as F

CHAIN = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]


def test_plan_by_image_and_by_rows():
  for G, world in ((32, 8), (8, 8), (16, 4), (2, 2), (1, 1)):
    seen = []
    for r in range(world):
      p = shard_plan(G, 2160, r, world)
      assert p["mode"] == "image" and p["rows"] == (0, 2160) and p["group"] == [r]
      seen += list(range(p["images"][0], p["images"][0] + p["images"][1]))
    assert seen == list(range(G))
  for G, world in ((1, 8), (2, 8), (4, 8), (1, 2), (2, 4)):
    cover = {}
    for r in range(world):
      p = shard_plan(G, 2160, r, world)
      assert p["mode"] == "rows" and p["images"][1] == 1 and r in p["group"] and len(p["group"]) == world // G
      cover.setdefault(p["images"][0], []).append(p["rows"])
    assert sorted(cover) == list(range(G))
    for spans in cover.values():
      spans.sort()
      assert spans[0][0] == 0 and sum(n for _, n in spans) == 2160
      assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1))
  with pytest.raises(ValueError):
    shard_plan(3, 2160, 0, 8)
  with pytest.raises(ValueError):
    shard_plan(1, 2161, 0, 8)


def _worker(rank, world, port, q):
  os.environ["MASTER_ADDR"] = "127.0.0.1"
  os.environ["MASTER_PORT"] = str(port)
  torch.set_num_threads(2)
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    H, W = 24, 20
    x = F.synth_images(1, H, W, seed=5).double()
    gy = torch.randn(1, H, W, 3, generator=torch.Generator().manual_seed(6), dtype=torch.float64)
    lgs = [F.synth_logits(f, 1, seed=7 + k).double() * 0.7 for k, f in enumerate(CHAIN)]
    p = shard_plan(1, H, rank, world)
    r0, n = p["rows"]
    y, gx, gl = F.chain_fwd_bwd(CHAIN, x[:, r0:r0 + n], lgs, gy[:, r0:r0 + n])
    flat = torch.cat([g.reshape(-1) for g in gl])
    dist.all_reduce(flat)                                   # the one exchange of the row-sharded path
    if rank == 0:
      yf, gxf, glf = F.chain_fwd_bwd(CHAIN, x, lgs, gy)
      ref = torch.cat([g.reshape(-1) for g in glf])
      q.put((float((flat - ref).abs().max() / ref.abs().max()), float((y - yf[:, r0:r0 + n]).abs().max()),
             float((gx - gxf[:, r0:r0 + n]).abs().max())))
  finally:
    dist.destroy_process_group()


def test_row_sharded_parameter_gradients_add_up():
  world = 2
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  port = 29700 + os.getpid() % 200
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  rel, ey, eg = q.get(timeout=180)
  for p in procs:
    p.join(timeout=60)
    assert p.exitcode == 0
  assert rel < 1e-12 and ey == 0.0 and eg == 0.0, (rel, ey, eg)
