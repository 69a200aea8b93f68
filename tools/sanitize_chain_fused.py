#!/usr/bin/env python
"""Tiny fused-chain launch set for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python tools/sanitize_chain_fused.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from exposure_b200.chain import FusedFilterChain  # noqa: E402

g = torch.Generator().manual_seed(0)
for (B, H, W, ids) in ((2, 64, 64, [0, 1, 2, 3, 4, 5, 6, 7]), (3, 33, 31, [7, 4, 9, 8]), (1, 40, 52, [3])):
  x = torch.rand(B, H, W, 3, generator=g).cuda()
  gy = torch.randn(B, H, W, 3, generator=g).cuda()
  ch = FusedFilterChain(ids, B, torch.device("cuda"))
  ch.logits.copy_((torch.randn(len(ids), B, 24, generator=g) * 0.5).cuda())
  for _ in range(2):
    y, gx, gl = ch.forward_backward(x, gy)
  torch.cuda.synchronize()
  print(B, H, W, ids, float(y.abs().sum()), float(gx.abs().sum()), float(gl.abs().sum()))
print("done")
