#!/usr/bin/env python
"""Randomised LIVE comparison of the oracle with the reference's own Python (build container only: needs
/root/reference): many random shapes / seeds / logit scales per filter, beyond the fixed vectors of
reference_golden.npz.  Run as a subprocess by tests/test_reference_live.py -- importing the TF stand-in
patches torch.Tensor (out-of-place `*=`), which must not leak into the test process.  Prints one JSON line."""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_reference_golden as G  # noqa: E402  (loads the shim + the reference modules; main() is not run)

sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import filters as OF  # noqa: E402
from oracle import nets as ON  # noqa: E402


def err(a, b, floor=1e-3):
  a, b = a.detach().double().reshape(-1), b.detach().double().reshape(-1)
  s = float(b.abs().max())
  if s == 0:
    return float(a.abs().max())
  return float(((a - b).abs() / b.abs().clamp_min(floor * s)).max())


def main():
  n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
  G.tf.set_float_dtype(torch.float64)
  rng = torch.Generator().manual_seed(2026)
  worst = {"param": 0.0, "y": 0.0, "gx": 0.0, "glogits": 0.0, "mask": 0.0, "masked": 0.0, "stats": 0.0}
  mcfg = G.util.Dict(dict(G.cfg))
  mcfg.masking = True
  for case in range(n_cases):
    fid = case % 10
    B = int(torch.randint(1, 4, (1,), generator=rng))
    H = int(torch.randint(1, 24, (1,), generator=rng))
    W = int(torch.randint(1, 24, (1,), generator=rng))
    scale = float(torch.rand(1, generator=rng)) * 2.5 + 0.1
    x = G.images(B, H, W, seed=9000 + case) if H * W * B * 3 >= 400 else \
        torch.exp(torch.randn(B, H, W, 3, generator=rng, dtype=torch.float64) - 1.6).clamp(0, 4)
    xr = x.clone().requires_grad_(True)
    f = G.quiet(G.FILTERS[fid], xr, G.cfg)
    n = f.get_num_filter_parameters()
    logits = (torch.randn(B, n, generator=rng, dtype=torch.float64) * scale).requires_grad_(True)
    gy = torch.randn(B, H, W, 3, generator=rng, dtype=torch.float64)
    param = f.filter_param_regressor(logits)
    y = f.process(xr, param)
    (y * gy).sum().backward()
    z = lambda g, like: torch.zeros_like(like) if g is None else g
    # oracle
    op = OF.regress(fid, logits.detach())
    oy = OF.process(fid, x, op)
    ogx, ogp = OF.process_bwd_analytic(fid, x, op, gy)
    ogl = OF.regress_bwd(fid, logits.detach(), ogp)
    worst["param"] = max(worst["param"], err(op, param.reshape(B, -1)))
    worst["y"] = max(worst["y"], err(oy, y))
    a, b = ogx, z(xr.grad, xr)
    if fid == OF.SP:        # away from exact channel ties (no RGBToHSV gradient in TF 1.6; tie conventions differ)
      px = x.clamp(max=1.0).reshape(-1, 3)
      keep = (px[:, 0] != px[:, 1]) & (px[:, 1] != px[:, 2]) & (px[:, 0] != px[:, 2])
      a, b = a.reshape(-1, 3)[keep], b.reshape(-1, 3)[keep]
    if a.numel():
      worst["gx"] = max(worst["gx"], err(a, b))
    worst["glogits"] = max(worst["glogits"], err(ogl, z(logits.grad, logits)))
    # Filter.get_mask + lerp with masking on, raw mask logits given directly
    fm = G.quiet(G.FILTERS[fid], x, mcfg)
    nm = fm.get_num_mask_parameters()
    ml = torch.randn(B, nm, generator=rng, dtype=torch.float64) * 0.8
    mask = G.quiet(fm.get_mask, x, ml)
    omask = OF.get_mask(fid, x, ml, True)
    worst["mask"] = max(worst["mask"], err(omask.expand_as(mask), mask))
    low = G.util.lerp(x, fm.process(x, param.detach()), mask)
    worst["masked"] = max(worst["masked"], err(OF.apply_masked(fid, x, logits.detach(), ml, True), low))
    # critic statistics (critics.py:48-76) through the reference critic's own code path is covered by the
    # fixture; here: the three statistics alone, restated inline from critics.py for arbitrary shapes
    if H * W >= 4:
      lum = (x[:, :, :, 0] * 0.27 + x[:, :, :, 1] * 0.67 + x[:, :, :, 2] * 0.06 + 1e-5)
      luminance, contrast = G.tf.nn.moments(lum, axes=[1, 2])
      i_max = G.tf.reduce_max(G.tf.clip_by_value(x, 0.0, 1.0), reduction_indices=[3])
      i_min = G.tf.reduce_min(G.tf.clip_by_value(x, 0.0, 1.0), reduction_indices=[3])
      sat = (i_max - i_min) / (G.tf.minimum(x=i_max + i_min, y=2.0 - i_max - i_min) + 1e-2)
      saturation, _ = G.tf.nn.moments(sat, axes=[1, 2])
      want = torch.stack([luminance, contrast, saturation], dim=1)
      worst["stats"] = max(worst["stats"], err(ON.critic_stats(x), want))
  print(json.dumps({"cases": n_cases, "worst": worst}))


if __name__ == "__main__":
  main()
