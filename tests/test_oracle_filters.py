"""CPU tests of the oracle itself (the oracle is test infrastructure): self-consistency --
fp32-vs-fp64 agreement, autograd-vs-closed-form gradients, finite differences, invariants.  The pin
against the reference's own Python is tests/test_reference_golden.py."""
import numpy as np
import pytest
import torch

from oracle import filters as F

ALL = list(range(8))


def _case(fid, B=3, H=12, W=10, dtype=torch.float64, seed=7):
  x = F.synth_images(B, H, W, seed=seed, dtype=dtype)
  lg = F.synth_logits(fid, B, seed=seed, dtype=dtype)
  return x, lg


@pytest.mark.parametrize("fid", ALL)
def test_fp32_tracks_fp64(fid):
  x, lg = _case(fid)
  y64 = F.process(fid, x, F.regress(fid, lg))
  y32 = F.process(fid, x.float(), F.regress(fid, lg.float())).double()
  # 2e-4: ContrastFilter's -cos(pi l)/2+1/2 and S+'s 1-s' cancel in fp32 (hard part 1)
  tol = 2e-4 if fid in (F.SP, F.CT) else 2e-6
  assert ((y32 - y64).abs() <= tol * y64.abs().clamp_min(1e-4)).all()


@pytest.mark.parametrize("fid", ALL)
def test_closed_form_grad_matches_autograd(fid):
  x, lg = _case(fid)
  p = F.regress(fid, lg)
  gy = torch.randn(x.shape, generator=torch.Generator().manual_seed(21), dtype=x.dtype)
  gxa, gpa = F.process_bwd_analytic(fid, x, p, gy)
  gxg, gpg = F.process_bwd_autograd(fid, x, p, gy)
  # S+: the restated TF constants 1/6, 2/6, 4/6 are fp32-rounded, which perturbs hue by 3e-8
  tol = 1e-6 if fid == F.SP else 1e-10
  assert (gxa - gxg).abs().max() <= tol * (1 + gxg.abs().max())
  assert (gpa - gpg).abs().max() <= tol * (1 + gpg.abs().max())


@pytest.mark.parametrize("fid", ALL)
def test_param_grad_finite_difference(fid):
  x, lg = _case(fid, B=2, H=6, W=5)
  p = F.regress(fid, lg)
  gy = torch.randn(x.shape, generator=torch.Generator().manual_seed(22), dtype=x.dtype)
  _, gp = F.process_bwd_analytic(fid, x, p, gy)
  eps = 1e-6
  for j in range(p.shape[1]):
    dp = torch.zeros_like(p)
    dp[:, j] = eps
    fd = ((F.process(fid, x, p + dp) - F.process(fid, x, p - dp)) * gy).sum(dim=(1, 2, 3)) / (2 * eps)
    assert torch.allclose(fd, gp[:, j], rtol=1e-5, atol=1e-7), (fid, j)


@pytest.mark.parametrize("fid", ALL)
def test_image_grad_finite_difference(fid):
  # smooth inputs only (no exact ties / knots): finite differences are meaningless on kinks
  g = torch.Generator().manual_seed(3)
  x = (torch.rand(2, 5, 4, 3, generator=g, dtype=torch.float64) * 0.9 + 0.03)
  lg = F.synth_logits(fid, 2, dtype=torch.float64)
  p = F.regress(fid, lg)
  gy = torch.randn(x.shape, generator=g, dtype=torch.float64)
  gx, _ = F.process_bwd_analytic(fid, x, p, gy)
  eps = 1e-7
  flat = x.reshape(-1)
  idx = torch.randperm(flat.numel(), generator=g)[:40]
  for i in idx.tolist():
    d = torch.zeros_like(flat)
    d[i] = eps
    fd = ((F.process(fid, (flat + d).reshape(x.shape), p) - F.process(fid, (flat - d).reshape(x.shape), p)) * gy).sum() / (2 * eps)
    ref = gx.reshape(-1)[i]
    assert abs(fd - ref) <= 2e-5 * (1 + abs(ref)), (fid, i, float(fd), float(ref))


def test_identity_at_default_parameters():
  """SURVEY 8c invariants: zero logits are the identity (up to each filter's clamps)."""
  x = F.synth_images(2, 8, 8, dtype=torch.float64)
  z = lambda fid: torch.zeros(2, F.NUM_PARAMS[fid], dtype=torch.float64)
  assert torch.allclose(F.apply_filter(F.E, x, z(F.E)), x)
  assert torch.allclose(F.apply_filter(F.G, x, z(F.G)), x.clamp_min(0.001))
  assert torch.allclose(F.apply_filter(F.W, x, z(F.W)), x / (1 + 1e-5))
  assert torch.allclose(F.apply_filter(F.T, x, z(F.T)), x.clamp(0, 1), atol=1e-12)
  assert torch.allclose(F.apply_filter(F.C, x, z(F.C)), x.clamp(0, 1), atol=1e-12)
  assert torch.allclose(F.apply_filter(F.CT, x, z(F.CT)), x)
  big = torch.full((2, 1), -60.0, dtype=torch.float64)     # sigmoid -> 0
  assert torch.allclose(F.apply_filter(F.BW, x, big), x)
  assert torch.allclose(F.apply_filter(F.SP, x, big), x.clamp_max(1.0))


def test_curves_monotone_and_saturate():
  xs = torch.linspace(-0.2, 1.3, 301, dtype=torch.float64).reshape(1, 1, 301, 1).repeat(1, 1, 1, 3)
  for fid in (F.T, F.C):
    lg = F.synth_logits(fid, 1, dtype=torch.float64)
    y = F.apply_filter(fid, xs, lg)
    assert (y[0, 0, 1:] - y[0, 0, :-1] >= -1e-12).all()
    assert torch.allclose(y[0, 0, -1], torch.ones(3, dtype=torch.float64))
    assert (y[0, 0, 0] == 0).all()


def test_rgb_hsv_round_trip():
  x = torch.rand(1, 16, 16, 3, dtype=torch.float64)
  h, s, v = F.rgb_to_hsv(x)
  assert torch.allclose(F.hsv_to_rgb(h, s, v), x, atol=1e-7)
  # known answers of the TF kernel definition: pure red / green / blue / grey
  px = torch.tensor([[[[1.0, 0, 0], [0, 1.0, 0], [0, 0, 1.0], [0.5, 0.5, 0.5]]]], dtype=torch.float64)
  h, s, v = F.rgb_to_hsv(px)
  assert torch.allclose(h[0, 0], torch.tensor([0.0, 1 / 3, 2 / 3, 0.0], dtype=torch.float64), atol=1e-7)
  assert torch.allclose(s[0, 0], torch.tensor([1.0, 1.0, 1.0, 0.0], dtype=torch.float64))
  assert torch.allclose(v[0, 0], torch.tensor([1.0, 1.0, 1.0, 0.5], dtype=torch.float64))


def test_regressor_ranges():
  f = torch.linspace(-20, 20, 41, dtype=torch.float64)[:, None]
  p = F.regress(F.E, f)
  assert p.min() >= -3.5 and p.max() <= 3.5 and abs(float(p[20])) < 1e-12
  g = F.regress(F.G, f)
  assert abs(float(g[20]) - 1) < 1e-12 and g.max() <= 3 + 1e-9 and g.min() >= 1 / 3 - 1e-9
  t = F.regress(F.T, f.repeat(1, 8))
  assert abs(float(t[20, 0]) - 1.25) < 1e-12
  c = F.regress(F.C, f.repeat(1, 24))
  assert abs(float(c[20, 0]) - 1.0) < 1e-9
  w = F.regress(F.W, f.repeat(1, 3))
  lum = 0.27 * w[:, 0] + 0.67 * w[:, 1] + 0.06 * w[:, 2]
  assert torch.allclose(lum, torch.ones_like(lum), atol=1e-4)


def test_chain_grad_matches_autograd_end_to_end():
  ids = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]
  B = 2
  x = F.synth_images(B, 6, 6, dtype=torch.float64, stress=False).requires_grad_(True)
  lgs = [F.synth_logits(f, B, dtype=torch.float64).requires_grad_(True) for f in ids]
  y = F.chain_fwd(ids, x, lgs)[-1]
  gout = torch.randn(y.shape, generator=torch.Generator().manual_seed(12), dtype=torch.float64)
  grads = torch.autograd.grad(y, [x] + lgs, grad_outputs=gout)
  _, gimg, glg = F.chain_fwd_bwd(ids, x.detach(), [l.detach() for l in lgs], gout)
  # 1e-5 of the largest entry: S+'s closed form vs the fp32-rounded TF hue constants (see above)
  assert (gimg - grads[0]).abs().max() <= 1e-5 * grads[0].abs().max()
  for a, b in zip(glg, grads[1:]):
    assert (a - b).abs().max() <= 1e-5 * b.abs().max() + 1e-12


def test_hsv_functors_agree_with_an_independent_implementation():
  """The RGB<->HSV functors are the one primitive restated from TF's kernel source
  (tensorflow/core/kernels/colorspace_op.h) rather than from the reference repo.  Python's colorsys is an
  independent implementation of the same hexcone model: 2000 random pixels + the tie / grey / black cases."""
  import colorsys
  g = torch.Generator().manual_seed(77)
  px = torch.rand(2000, 3, generator=g, dtype=torch.float64)
  px[0] = 0.0
  px[1] = 0.37
  px[2] = torch.tensor([0.5, 0.5, 0.2])
  px[3] = torch.tensor([0.2, 0.7, 0.7])
  px[4] = torch.tensor([0.9, 0.1, 0.9])
  h, s, v = F.rgb_to_hsv(px.reshape(1, 1, -1, 3))
  h, s, v = h.reshape(-1), s.reshape(-1), v.reshape(-1)
  back = F.hsv_to_rgb(h.reshape(1, 1, -1, 1), s.reshape(1, 1, -1, 1), v.reshape(1, 1, -1, 1)).reshape(-1, 3)
  for i in range(px.shape[0]):
    r, gg, b = (float(t) for t in px[i])
    hh, ss, vv = colorsys.rgb_to_hsv(r, gg, b)
    assert abs(float(s[i]) - ss) < 1e-12 and abs(float(v[i]) - vv) < 1e-12
    dh = abs(float(h[i]) - hh)
    assert min(dh, 1 - dh) < 1e-7, (i, float(h[i]), hh)     # float32(1/6, 2/6, 4/6) constants, like TF's float kernel; hue is circular
    rr, g2, bb = colorsys.hsv_to_rgb(float(h[i]), float(s[i]), float(v[i]))
    assert max(abs(float(back[i, 0]) - rr), abs(float(back[i, 1]) - g2), abs(float(back[i, 2]) - bb)) < 1e-6


def _tie_pixels():
  """SaturationPlus inputs with exact channel ties (two channels share the max or the min), both sides of V = 0.5."""
  px = []
  for hi, lo in ((0.8, 0.3), (0.4, 0.1), (0.9, 0.6), (0.3, 0.05)):
    px += [(hi, hi, lo), (hi, lo, hi), (lo, hi, hi),        # tie for the max
           (hi, lo, lo), (lo, hi, lo), (lo, lo, hi)]        # tie for the min
  return torch.tensor(px, dtype=torch.float64).reshape(1, 1, -1, 3)


def test_satplus_tie_gradient_is_a_one_sided_derivative_of_the_forward():
  """VERDICT r1 weak #3.  TF 1.6 registers no gradient for RGBToHSV, so the reference defines none for
  SaturationPlusFilter's image input; where two channels tie for the max or the min the forward has a kink
  and the Jacobian is not unique.  The rule the oracle and the CUDA kernels implement -- the FIRST channel in R, G, B
  order that attains the max (min) is treated as THE max (min) -- is pinned here to the reference-pinned FORWARD: every
  entry dy_o/dx_c it produces equals the right or the left derivative of the forward along x_c (whichever side keeps
  the ordering the rule assumed).  (Grey pixels r = g = b are a DISCONTINUITY of TF's HSV round trip -- hue jumps
  between 0, 1/3, 2/3 -- so no derivative exists there; the kernels use the hue-0 branch TF's forward takes.)"""
  x = _tie_pixels()
  p = torch.full((1, 1), 0.7, dtype=torch.float64)
  # step: large against the 1e-8 wobble the float32-rounded TF constants (1/6, 2/6, 4/6) put on the forward exactly at a
  # tie, small against its curvature
  h = 1e-4
  fwd = lambda t: F.process(F.SP, t, p)
  y0 = fwd(x)
  for o in range(3):
    gy = torch.zeros_like(x)
    gy[..., o] = 1.0
    J_o, _ = F.process_bwd_analytic(F.SP, x, p, gy)                       # row o of the Jacobian: dy_o / dx_c
    for c in range(3):
      d = torch.zeros_like(x)
      d[..., c] = h
      right = (fwd(x + d) - y0)[..., o] / h
      left = (y0 - fwd(x - d))[..., o] / h
      a = J_o[..., c]
      ok = ((a - right).abs() <= 2e-3 * (1 + right.abs())) | ((a - left).abs() <= 2e-3 * (1 + left.abs()))
      assert bool(ok.all()), (o, c, a[~ok].tolist(), right[~ok].tolist(), left[~ok].tolist())
  # and the kink is real: at these pixels the two sides differ for some entry (the test above is not vacuous)
  d = torch.zeros_like(x)
  d[..., 0] = h
  assert float(((fwd(x + d) - y0) / h - (y0 - fwd(x - d)) / h).abs().max()) > 1e-2
