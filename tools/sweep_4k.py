"""BASELINE configs[4]: 4K (3840 x 2160) 8-filter chain fwd+bwd, global batch 1 -> 32 frames, on the N GPUs this
process group spans (run under torchrun for N > 1; plain python for N = 1).  One JSON line per batch size:
frames/s (whole job, CUDA events, max over ranks), per-GPU HBM-roofline fractions, sharding mode.  With fewer frames
than GPUs the frames are sharded by ROWS and the parameter gradients completed by one small all-reduce
(exposure_b200.chain.ShardedFilterChain); the row-sharded result is checked against a single-GPU run of the frame."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exposure_b200.chain import FusedFilterChain, ShardedFilterChain   # noqa: E402
from oracle import filters as F                                        # noqa: E402

CHAIN = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]
H, W = 2160, 3840


def block(img, r0, rows, dev, what):
  """rows [r0, r0 + rows) of synthetic frame `img` (any rank can regenerate any block: seeded per 270-row stripe)."""
  out = []
  for s0 in range(r0, r0 + rows, 270):
    g = torch.Generator(device=dev).manual_seed(1000003 * img + s0 + (7 if what == "gy" else 0))
    n = min(270, r0 + rows - s0)
    z = torch.randn(n, W, 3, device=dev, generator=g)
    out.append(torch.exp(z - 3.2).clamp_(0, 4) if what == "x" else z)
  return torch.cat(out)


def main():
  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  peak = 6538.0
  pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
  if os.path.exists(pk):
    peak = float(json.load(open(pk))["hbm_gbs"])
  steps = int(os.environ.get("SWEEP_STEPS", "10"))
  for G in (1, 2, 4, 8, 16, 32):
    if G >= world and G % world:
      continue
    ch = ShardedFilterChain(CHAIN, G, H, W, dev, rank, world)
    lgs = [(F.synth_logits(f, G, seed=50 + k) * 0.7).to(dev) for k, f in enumerate(CHAIN)]
    ch.set_logits(lgs)
    b0, nb = ch.plan["images"]
    r0, rows = ch.plan["rows"]
    x = torch.stack([block(b0 + i, r0, rows, dev, "x") for i in range(nb)])
    gy = torch.stack([block(b0 + i, r0, rows, dev, "gy") for i in range(nb)])
    y, gx = torch.empty_like(x), torch.empty_like(x)
    for _ in range(3):
      _, _, gl = ch.forward_backward(x, gy, y_out=y, gx_out=gx)
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
      _, _, gl = ch.forward_backward(x, gy, y_out=y, gx_out=gx)
    e1.record()
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / steps
    check = None
    if ch.plan["mode"] == "rows" and rank == 0:
      # the frame rank 0 shares, recomputed whole on this GPU: y / gx rows bit-identical, parameter gradients to
      # reduction order
      xf, gf = block(0, 0, H, dev, "x")[None], block(0, 0, H, dev, "gy")[None]
      ref = FusedFilterChain(CHAIN, 1, dev)
      ref.set_logits([l[0:1] for l in lgs])
      yr, gxr, glr = ref.forward_backward(xf, gf)
      torch.cuda.synchronize()
      rel = float((gl[:, 0] - glr[:, 0]).abs().max() / glr.abs().max())
      check = {"rows_bit_identical": bool(torch.equal(y[0], yr[0, r0:r0 + rows]) and torch.equal(gx[0], gxr[0, r0:r0 + rows])),
               "param_grad_rel_dev": rel}
      del xf, gf, yr, gxr
    if rank == 0:
      px_gpu = nb * rows * W                                  # pixels one GPU processes per step
      print(json.dumps({"workload": "chain8 fwd+bwd, 4K frames (3840x2160x3 fp32)", "n_gpus": world, "global_batch": G,
                        "sharding": ch.plan["mode"], "frames_per_gpu": nb, "rows_per_gpu": rows, "ms_per_step": ms,
                        "frames_per_s": G / ms * 1e3,
                        "per_gpu_algorithmic_gbs_60N": px_gpu * 480 / ms / 1e6, "per_gpu_frac_60N": px_gpu * 480 / ms / 1e6 / peak,
                        "per_gpu_hbm_bytes_moved_gbs": px_gpu * 48 / ms / 1e6, "per_gpu_hbm_frac_moved": px_gpu * 48 / ms / 1e6 / peak,
                        "hbm_peak_gbs": peak, "steps": steps, "row_shard_check": check}), flush=True)
    del ch, x, gy, y, gx
    torch.cuda.empty_cache()
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
