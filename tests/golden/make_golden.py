#!/usr/bin/env python
"""Generates tests/golden/filters_golden.npz from the oracle (fp64), seeded.

These vectors come from the ORACLE (oracle/filters.py), not from the reference: they guard the oracle
against drift and give the CUDA kernels a fixed target that does not depend on re-running it.  The
vectors that come from executing the reference's own Python are tests/golden/reference_golden.npz
(make_reference_golden.py).

  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import filters as F  # noqa: E402


def main():
  B, H, W = 3, 12, 10
  out = {}
  for fid in range(8):
    x = F.synth_images(B, H, W, seed=100 + fid, dtype=torch.float64)
    lg = F.synth_logits(fid, B, seed=200, dtype=torch.float64)
    p = F.regress(fid, lg)
    y = F.process(fid, x, p)
    gy = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(300 + fid), dtype=torch.float64)
    gx, gp = F.process_bwd_analytic(fid, x, p, gy)
    gl = F.regress_bwd(fid, lg, gp)
    for k, v in dict(x=x, logits=lg, params=p, y=y, gy=gy, gx=gx, gparams=gp, glogits=gl).items():
      out["f%d_%s" % (fid, k)] = v.numpy()
  ids = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]
  x = F.synth_images(2, 16, 16, seed=77, dtype=torch.float64)
  lgs = [F.synth_logits(f, 2, seed=78, dtype=torch.float64) * 0.5 for f in ids]
  gout = torch.randn(2, 16, 16, 3, generator=torch.Generator().manual_seed(79), dtype=torch.float64)
  y, gimg, glg = F.chain_fwd_bwd(ids, x, lgs, gout)
  out.update(chain_x=x.numpy(), chain_gout=gout.numpy(), chain_y=y.numpy(), chain_gx=gimg.numpy())
  for k, (l, g) in enumerate(zip(lgs, glg)):
    out["chain_logits%d" % k] = l.numpy()
    out["chain_glogits%d" % k] = g.numpy()
  path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "filters_golden.npz")
  np.savez_compressed(path, **out)
  print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
  main()
