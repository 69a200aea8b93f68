#!/usr/bin/env python
"""Isolated device time of single ops at the train step's shapes: CUDA-graph replay of 20 back-to-back launches, CUDA
events around the replay (no launch gaps, no concurrency with other kernels -- what a CUPTI timeline of the whole
iteration cannot tell).  Usage: python tools/op_bench.py [substring ...]"""
import sys

import torch

sys.path.insert(0, ".")
from exposure_b200 import nn_ops as K  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
R = lambda *s: torch.randn(*s, device=dev, generator=g)
B = 64
ops = {}


def first_layer(tag, n, Cv):
  x, vec = R(n, 64, 64, 3).abs() * 0.3, R(n, Cv)
  W, b = R(4, 4, 3 + Cv, 32) * 0.05, R(32) * 0.1
  dy = R(n, 32, 32, 32)
  gW = torch.zeros(4, 4, 3 + Cv, 32, device=dev)
  y = K.conv_fwd(x, W, b, vec=vec, shift=0.5)
  ops["first_fwd_%s" % tag] = lambda: K.conv_fwd(x, W, b, vec=vec, shift=0.5, out=y)
  ops["first_wgrad_%s" % tag] = lambda: K.conv_wgrad(x, dy, vec=vec, shift=0.5, out=gW)
  ops["first_dgrad_%s" % tag] = lambda: K.conv_first_dgrad(dy, W, Cv, (64, 64))
  ops["first_dgrad_imgonly_%s" % tag] = lambda: K.conv_first_dgrad(dy, W, Cv, (64, 64), need_vec=False)
  ops["first_dgrad_veconly_%s" % tag] = lambda: K.conv_first_dgrad(dy, W, Cv, (64, 64), need_image=False)


first_layer("policy_b64_cv11", B, 11)
first_layer("critic_b192_cv3", 3 * B, 3)
first_layer("critic_b64_cv3", B, 3)

for name, (IH, Cin, Cout) in {"l2": (32, 32, 64), "l3": (16, 64, 128), "l4": (8, 128, 256)}.items():
  for n in (B, 3 * B):
    x, W, b, dy = R(n, IH, IH, Cin), R(4, 4, Cin, Cout) * 0.05, R(Cout) * 0.1, R(n, IH // 2, IH // 2, Cout)
    y = K.conv_fwd(x, W, b)
    dx = torch.empty_like(x)
    gW = torch.zeros_like(W)
    ops["conv_fwd_%s_b%d" % (name, n)] = (lambda x=x, W=W, b=b, y=y: K.conv_fwd(x, W, b, out=y))
    ops["conv_dgrad_%s_b%d" % (name, n)] = (lambda x=x, W=W, dy=dy, dx=dx: K.conv_dgrad(dy, W, tuple(x.shape), a_in=x, out=dx))
    ops["conv_wgrad_%s_b%d" % (name, n)] = (lambda x=x, dy=dy, gW=gW: K.conv_wgrad(x, dy, out=gW))

img = R(B, 64, 64, 3).abs() * 0.3
st = K.stats_fwd(img)
ops["stats_fwd_b64"] = lambda: K.stats_fwd(img)
u = R(B, 64, 64, 3)
ops["stats_jvp_b64"] = lambda: K.stats_jvp(img, st, u)
gs = R(B, 3)
ops["stats_bwd_b64"] = lambda: K.stats_bwd(img, st, gs, g_direct=u)
h, Wf, bf = R(B, 4096), R(4096, 128) * 0.02, R(128)
dh = R(B, 128)
gWf = torch.zeros_like(Wf)
ops["fc_fwd_4096x128_b64"] = lambda: K.fc_fwd(h, Wf, bf)
ops["fc_dgrad_4096x128_b64"] = lambda: K.fc_dgrad(dh, Wf)
ops["fc_wgrad_4096x128_b64"] = lambda: K.fc_wgrad(h, dh, out=gWf)

sel = [a for a in sys.argv[1:]]
for name, fn in ops.items():
  if sel and not any(s in name for s in sel):
    continue
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  gr = torch.cuda.CUDAGraph()
  with torch.cuda.graph(gr):
    for _ in range(20):
      fn()
  gr.replay()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(5):
    gr.replay()
  e1.record()
  torch.cuda.synchronize()
  print("%-36s %8.2f us" % (name, e0.elapsed_time(e1) * 10.0))
