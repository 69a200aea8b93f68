"""Deterministic network weights by variable name -- shared by tests/golden/make_reference_golden.py
(which runs the reference's Python with them) and the parity tests (which load the same values into
the oracle / the CUDA trainer).  The shipped pretrained checkpoint cannot travel to the GPU box, these
can: every value is a pure function of (seed, variable name, shape).

Kernels / matrices: xavier-uniform like the reference's initialisers (ly.conv2d / ly.fully_connected
defaults, filters.py:37,43) -- U(-l, l), l = sqrt(6 / (fan_in + fan_out)); biases: 0.05 * N(0,1) instead
of zeros so that they matter."""
import math

import torch


def _name_hash(s):
  h = 0
  for c in s.encode():
    h = (h * 131 + c) % 2147483647
  return h


def make(name, shape, seed=0, dtype=torch.float64):
  g = torch.Generator().manual_seed(seed * 1000003 + _name_hash(name) % 1000003)
  shape = tuple(int(s) for s in shape)
  if len(shape) == 1:
    return (0.05 * torch.randn(shape, generator=g, dtype=torch.float64)).to(dtype)
  if len(shape) == 4:
    k2 = shape[0] * shape[1]
    fan_in, fan_out = k2 * shape[2], k2 * shape[3]
  else:
    fan_in, fan_out = shape
  lim = math.sqrt(6.0 / (fan_in + fan_out))
  return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim).to(dtype)


def make_named(names_and_shapes, seed=0, dtype=torch.float64):
  return {n: make(n, s, seed, dtype) for n, s in names_and_shapes.items()}
