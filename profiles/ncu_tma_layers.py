#!/usr/bin/env python
"""One eager launch of each TMA-fed tcgen05 conv kernel type at the batch-64 layer shapes, for
  ncu --set full --clock-control none --import-source on -k "regex:tma_gemm|conv_first|conv_dgrad_img3" -o gpurun_out/prof_tma python profiles/ncu_tma_layers.py
"""
import sys

import torch

sys.path.insert(0, ".")
from exposure_b200 import nn_ops as K  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
B = 64
for IH, Cin, Cout in [(64, 14, 32), (32, 32, 64), (16, 64, 128), (8, 128, 256)]:
  x = torch.randn(B, IH, IH, Cin, device=dev, generator=g)
  W = torch.randn(4, 4, Cin, Cout, device=dev, generator=g) * 0.05
  b = torch.zeros(Cout, device=dev)
  dy = torch.randn(B, IH // 2, IH // 2, Cout, device=dev, generator=g)
  for _ in range(2):                      # second launch of each is the warm one
    K.conv_fwd(x, W, b)
    if Cin % 32 == 0:
      K.conv_dgrad(dy, W, tuple(x.shape), a_in=x)
    K.conv_wgrad(x, dy)
# the split first layer (csrc/conv_first.cu): policy (11 state channels) and critic (3 statistic channels) inputs
for Cv in (11, 3):
  x = torch.rand(B, 64, 64, 3, device=dev, generator=g) * 0.3
  vec = torch.randn(B, Cv, device=dev, generator=g)
  W = torch.randn(4, 4, 3 + Cv, 32, device=dev, generator=g) * 0.05
  b = torch.zeros(32, device=dev)
  dy = torch.randn(B, 32, 32, 32, device=dev, generator=g)
  for _ in range(2):
    K.conv_fwd(x, W, b, vec=vec, shift=0.5)
    K.conv_wgrad(x, dy, vec=vec, shift=0.5)
    K.conv_first_dgrad(dy, W, Cv, (64, 64))
torch.cuda.synchronize()
