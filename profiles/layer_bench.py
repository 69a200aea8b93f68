#!/usr/bin/env python
"""Per-layer timing of the conv primitives (fprop / dgrad / wgrad) for every GEMM backend.

  python profiles/layer_bench.py [--batch 64] [--backends 1,2,3] > gpurun_out/layer_bench.md

Times (CUDA-graph replay, device clock) each layer shape of the policy / value / critic stacks (agent.py:12-41, 64x64 input, 4x4
stride-2 convolutions, channels 32-64-128-256) with CUDA events on the launching stream, inputs
rotated through buffers larger than L2 is NOT attempted here: the activations of one layer are a
few MB and live in L2 in the training step too, so the warm numbers are the relevant ones.
"""
import argparse
import sys

import torch

sys.path.insert(0, ".")
from exposure_b200 import nn_ops as K  # noqa: E402


def timed(fn, iters=20, reps=5):
  """us per call, device time of a CUDA graph holding `iters` calls (no host launch overhead)."""
  for _ in range(3):
    fn()
  torch.cuda.synchronize()
  g = torch.cuda.CUDAGraph()
  side = torch.cuda.Stream()
  side.wait_stream(torch.cuda.current_stream())
  with torch.cuda.stream(side):
    with torch.cuda.graph(g, stream=side):
      for _ in range(iters):
        fn()
  torch.cuda.current_stream().wait_stream(side)
  g.replay()
  torch.cuda.synchronize()
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(reps):
    g.replay()
  e1.record()
  torch.cuda.synchronize()
  return e0.elapsed_time(e1) / (iters * reps) * 1e3   # us


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--batch", type=int, default=64)
  ap.add_argument("--backends", default="1,2,3,4")
  a = ap.parse_args()
  dev = torch.device("cuda:0")
  g = torch.Generator(device=dev).manual_seed(1)
  B = a.batch
  layers = [(64, 14, 32), (32, 32, 64), (16, 64, 128), (8, 128, 256)]   # (IH, Cin, Cout)
  names = {1: "cuda-cores", 2: "tcgen05", 3: "tcgen05-ws", 4: "tcgen05-tma"}
  print("| layer (IH,Cin,Cout) | GFLOP | backend | fprop us (TF/s) | dgrad us (TF/s) | wgrad us (TF/s) |")
  print("|---|---:|---|---:|---:|---:|")
  for IH, Cin, Cout in layers:
    x = torch.randn(B, IH, IH, Cin, device=dev, generator=g)
    W = torch.randn(4, 4, Cin, Cout, device=dev, generator=g) * 0.05
    b = torch.zeros(Cout, device=dev)
    dy = torch.randn(B, IH // 2, IH // 2, Cout, device=dev, generator=g)
    y = torch.empty_like(dy)
    dx = torch.empty_like(x)
    gW = torch.empty_like(W)
    flop = 2.0 * B * (IH // 2) ** 2 * 16 * Cin * Cout
    for be in [int(s) for s in a.backends.split(",")]:
      K.set_gemm_backend(be)
      tf = timed(lambda: K.conv_fwd(x, W, b, out=y))
      td = timed(lambda: K.conv_dgrad(dy, W, tuple(x.shape), a_in=x, out=dx))
      tw = timed(lambda: K.conv_wgrad(x, dy, out=gW))
      f = lambda t: "%.1f (%.1f)" % (t, flop / t / 1e6)
      print("| %d,%d,%d | %.2f | %s | %s | %s | %s |" % (IH, Cin, Cout, flop / 1e9, names[be], f(tf), f(td), f(tw)))
  K.set_gemm_backend(0)


if __name__ == "__main__":
  main()
