#!/usr/bin/env python
"""Where does a K step of the persistent TMA-fed tcgen05 conv kernel spend its time?

Needs a development build with SM-clock stamps in the kernel:

  EXPOSURE_NVCC_EXTRA=-DEXPO_TMA_TRACE python -m exposure_b200.build
  gpurun -- 'EXPOSURE_NVCC_EXTRA=-DEXPO_TMA_TRACE python tools/tma_trace.py'

Prints, for the first CTAs of one warm launch of the layer-1 forward convolution at batch 64 (512 tiles of 8 K steps
over 148 CTAs), the stamps of every ring hand-over relative to the CTA's first one, in SM clocks:
  P   producer: slot free, TMA about to be issued        C0  converters: raw tile landed
  C1  converters: lo tiles written, arrived              M0  MMA issuer: operands ready
  M1  MMA issuer: 12 MMAs + commit issued
and per tile the epilogue's wait / drain."""
import ctypes
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from exposure_b200 import _cabi, nn_ops as K  # noqa: E402

CTAS, STEPS, TILES = 4, 96, 16
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
B = 64
which = sys.argv[1] if len(sys.argv) > 1 else "dgrad2"
LAYERS = {"2": (32, 32, 64), "3": (16, 64, 128), "4": (8, 128, 256)}        # IH, Cin, Cout
kind, layer = which[:-1], which[-1]
IH, Cin, Cout = LAYERS[layer]
x = torch.randn(B, IH, IH, Cin, device=dev, generator=g)
W = torch.randn(4, 4, Cin, Cout, device=dev, generator=g) * 0.05
b = torch.zeros(Cout, device=dev)
dy = torch.randn(B, IH // 2, IH // 2, Cout, device=dev, generator=g)
if kind == "fprop":
  run = lambda: K.conv_fwd(x, W, b)
elif kind == "dgrad":
  run = lambda: K.conv_dgrad(dy, W, tuple(x.shape), a_in=x)
else:
  run = lambda: K.conv_wgrad(x, dy)
for _ in range(3):
  run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print("%s: launch %.1f us (eager, CUDA events)" % (which, e0.elapsed_time(e1) * 1e3))
steps = np.zeros((CTAS, STEPS, 5), dtype=np.int64)
tiles = np.zeros((CTAS, TILES, 3), dtype=np.int64)
fn = _cabi.lib().exp_debug_tma_trace
fn.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
fn.restype = ctypes.c_int
assert fn(steps.ctypes.data, tiles.ctypes.data) == 0
cta = np.zeros((CTAS, 8), dtype=np.int64)
fn2 = _cabi.lib().exp_debug_tma_trace_cta
fn2.argtypes = [ctypes.c_void_p]
fn2.restype = ctypes.c_int
assert fn2(cta.ctypes.data) == 0
print("one-tile kernel, per CTA (SM clocks since entry): set-up done, accumulator ready, tile stored / parked, after cluster exchange, exit")
for c in range(CTAS):
  print("   CTA %d: %s   first TMA issue at %d" % (c, " ".join("%7d" % (cta[c, i] - cta[c, 0]) for i in range(1, 6)), steps[c, 0, 0] - cta[c, 0]))
  print("          first group of 4 float4 of the store pass: read at %d, stored at %d" % (cta[c, 6] - cta[c, 0], cta[c, 7] - cta[c, 0]))
for c in range(2):
  t0 = steps[c, 0, 0]
  print("CTA %d   it:     P     C0     C1     M0     M1   | dP  (SM clocks since the CTA's first TMA issue)" % c)
  prev = t0
  for it in range(40):
    r = steps[c, it] - t0
    print("      %4d: %6d %6d %6d %6d %6d | %5d" % (it, r[0], r[1], r[2], r[3], r[4], steps[c, it, 0] - prev))
    prev = steps[c, it, 0]
  print("   tiles: wait-begin, accumulator ready, drained")
  for j in range(5):
    r = tiles[c, j] - t0
    print("      %4d: %6d %6d %6d" % (j, r[0], r[1], r[2]))
d = np.diff(steps[:, 8:40, 0], axis=1)
print("steady state: %.0f clocks per K step (median over CTAs 0-3, steps 8-40)" % np.median(d))
