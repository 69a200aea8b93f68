"""S+ forward without the hue (csrc/filter_math.cuh satplus_full): whichever branch TF's rgb_to_hsv takes, the three ramps
hsv_to_rgb builds from the hue are d_c = (c - min) / (max - min); grey pixels keep hue 0, i.e. d = (1, 0, 0).  Checked
in float64 against the op-by-op restatement of tensorflow/core/kernels/colorspace_op.h in oracle/filters.py
(filters.py:484-498), incl. channel ties, pure greys, clipped highlights and zeros."""
import torch

from oracle import filters as F


def _closed_form(x, p):
  xm = x.clamp(max=1.0)
  V, m = xm.amax(dim=-1, keepdim=True), xm.amin(dim=-1, keepdim=True)
  rng = V - m
  col = rng > 0
  S = torch.where(col & (V > 0), rng / V.clamp_min(1e-300), torch.zeros_like(V))
  kk = 0.5 - (0.5 - V).abs()
  s2 = S + (1 - S) * kk * 0.8
  d = torch.where(col, (xm - m) / torch.where(col, rng, torch.ones_like(rng)), torch.tensor([1.0, 0.0, 0.0], dtype=x.dtype).expand_as(xm))
  full = ((1 - s2) + s2 * d) * V
  pp = p.reshape(-1, 1, 1, 1)
  return xm * (1 - pp) + full * pp


def test_closed_form_equals_hsv_round_trip():
  B, H, W = 3, 24, 24
  x = F.synth_images(B, H, W, seed=77).double()
  # exact ties and greys on top of the seeded stress values
  x[0, 0, :8] = torch.tensor([0.3, 0.3, 0.1], dtype=torch.float64)      # two channels tied at the maximum
  x[0, 1, :8] = torch.tensor([0.1, 0.4, 0.1], dtype=torch.float64)      # tied at the minimum
  x[0, 2, :8] = torch.tensor([0.25, 0.25, 0.25], dtype=torch.float64)   # grey
  x[0, 3, :8] = torch.tensor([2.0, 1.5, 0.2], dtype=torch.float64)      # clipped highlights tie at 1
  x[0, 4, :8] = 0.0
  p = torch.sigmoid(torch.randn(B, 1, generator=torch.Generator().manual_seed(5), dtype=torch.float64))
  ref = F.process(F.SP, x, p)
  got = _closed_form(x, p)
  # the restatement keeps TF's float32 constants 1/6, 2/6, 4/6 even in float64, which moves its ramps by up to ~5e-8:
  # that, not the closed form, is the whole difference
  assert torch.allclose(got, ref, rtol=0, atol=1e-7), float((got - ref).abs().max())
