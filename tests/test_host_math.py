"""CPU: the product's per-pixel device math (csrc/filter_math.cuh, filter_mask.cuh), compiled for
the host by tests/host_math/px_harness.cpp, against the CPU oracle.

This checks the closed-form forward / backward formulas, the per-image constant setup
(fused filter_param_regressor, curve prefix sums, slope table), the final transforms of the
reduced sums, the Level / Vignet filters and the spatial mask without a GPU.  The CUDA kernels
that tile these functions over a batch are covered by the `-m gpu` tests through the C ABI.

Tolerances: forward 2e-6 relative to max(|ref32|, 1e-4) (host libm vs oracle libm: same machine,
so only op-order differences remain; Contrast gets the cos cancellation allowance of
tests/test_filters_gpu.py); gradients 1e-4 of the per-image scale, against the fp64 oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import filters as F

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host_math", "px_harness.cpp")
PS = 24
fp = ctypes.POINTER(ctypes.c_float)


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
  out = str(tmp_path_factory.mktemp("hm") / "libpx_harness.so")
  subprocess.run(["g++", "-O1", "-ffp-contract=off", "-pthread", "-shared", "-fPIC", "-o", out, SRC], check=True)
  return ctypes.CDLL(out)


def _p(a):
  return None if a is None else a.ctypes.data_as(fp)


def _np(t):
  return np.ascontiguousarray(t.detach().cpu().numpy(), dtype=np.float32)


def _pad(lg):
  out = np.zeros((lg.shape[0], PS), np.float32)
  out[:, :lg.shape[1]] = _np(lg)
  return out


def _fwd_tol(fid, x, p, ref32):
  tol = 2e-6 * ref32.abs().clamp_min(1e-4)
  if fid == F.CT:
    lum = F.rgb2lum(x).clamp(0, 1)
    tol = tol + p.abs()[:, :, None, None] * x.abs() / (lum + 1e-6) * (2 * 2.0 ** -24)
  if fid == F.SP:
    # the fp32 reference turns a hue back into three ramps (2 - |6H - 2| ...): 6H carries up to ~6e-7 of ABSOLUTE rounding
    # error, i.e. up to 6e-7 * V * p in y.  The kernel computes the ramps in closed form ((c - min) / range, exact to ~1 ulp),
    # so it may differ from the fp32 restatement by that much while being closer to the fp64 one (checked below / by the bwd)
    V = x.clamp(max=1.0).amax(dim=-1, keepdim=True).abs()
    tol = tol + 1e-6 * V * p.abs().reshape(-1, 1, 1, 1)
  if fid == F.G:
    tol = tol + 4e-6 * ref32.abs()          # exp2(g*log2 x) vs powf
  return tol


@pytest.mark.parametrize("fid", range(10))
def test_forward_math(hm, fid):
  B, H, W = 3, 17, 23
  x = F.synth_images(B, H, W, seed=40 + fid)
  lg = F.synth_logits(fid, B)
  p32 = F.regress(fid, lg)
  ref = F.process(fid, x, p32)
  y = np.empty((B, H, W, 3), np.float32)
  xs, lgs = _np(x), _pad(lg)
  assert hm.hm_fwd(fid, _p(xs), _p(y), _p(lgs), PS, B, H * W, 1) == 0
  err = (torch.from_numpy(y) - ref).abs()
  tol = _fwd_tol(fid, x, p32, ref)
  assert (err <= tol).all(), "fid %d max err/tol %.3g" % (fid, float((err / tol).max()))


@pytest.mark.parametrize("fid", range(10))
def test_backward_math(hm, fid):
  B, H, W = 3, 17, 23
  x = F.synth_images(B, H, W, seed=50 + fid)
  lg = F.synth_logits(fid, B)
  gy = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(7 + fid))
  p64 = F.regress(fid, lg.double())
  gx64, gp64 = F.process_bwd_analytic(fid, x.double(), p64, gy.double())
  gl64 = F.regress_bwd(fid, lg.double(), gp64)
  gx = np.empty((B, H, W, 3), np.float32)
  gl = np.zeros((B, PS), np.float32)
  xs, gys, lgs = _np(x), _np(gy), _pad(lg)
  assert hm.hm_bwd(fid, _p(xs), _p(gys), _p(gx), _p(gl), _p(lgs), PS, B, H * W, 1) == 0
  e = (torch.from_numpy(gx).double() - gx64).abs()
  assert (e <= 1e-4 * gx64.abs().clamp_min(1e-3 * float(gx64.abs().max()) + 1e-30)).all(), float(e.max())
  n = F.NUM_PARAMS[fid]
  eg = (torch.from_numpy(gl[:, :n]).double() - gl64).abs()
  scale = gl64.abs().max(dim=1, keepdim=True).values.clamp_min(1e-6)
  assert (eg <= 2e-4 * scale).all(), (fid, float((eg / scale).max()))


@pytest.mark.parametrize("fid", [F.E, F.G, F.T, F.C])
def test_cfg_driven_ranges_math(hm, fid, monkeypatch):
  """cfg.exposure_range / gamma_range / tone_curve_range / color_curve_range (filters.py:179, 202, 261, 309) are
  launch arguments of the kernels, not baked constants: non-default ranges -- incl. a colour range that is NOT
  centred on its initial value 1, so util.tanh_range's bias is non-zero -- forward and backward vs the oracle."""
  rng = dict(exposure_range=2.0, gamma_range=2.5, tone=(0.25, 3.0), color=(0.8, 1.3))
  monkeypatch.setattr(F, "EXPOSURE_RANGE", rng["exposure_range"])
  monkeypatch.setattr(F, "GAMMA_RANGE", rng["gamma_range"])
  monkeypatch.setattr(F, "TONE_CURVE_RANGE", rng["tone"])
  monkeypatch.setattr(F, "COLOR_CURVE_RANGE", rng["color"])
  hm.hm_set_ranges.argtypes = [ctypes.c_float] * 6
  hm.hm_set_ranges(rng["exposure_range"], rng["gamma_range"], *rng["tone"], *rng["color"])
  try:
    B, H, W = 3, 13, 11
    x = F.synth_images(B, H, W, seed=70 + fid)
    lg = F.synth_logits(fid, B)
    gy = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(17 + fid))
    p32 = F.regress(fid, lg)
    ref = F.process(fid, x, p32)
    y = np.empty((B, H, W, 3), np.float32)
    xs, gys, lgs = _np(x), _np(gy), _pad(lg)
    assert hm.hm_fwd(fid, _p(xs), _p(y), _p(lgs), PS, B, H * W, 1) == 0
    err = (torch.from_numpy(y) - ref).abs()
    tol = _fwd_tol(fid, x, p32, ref)
    assert (err <= tol).all(), "fid %d max err/tol %.3g" % (fid, float((err / tol).max()))
    p64 = F.regress(fid, lg.double())
    gx64, gp64 = F.process_bwd_analytic(fid, x.double(), p64, gy.double())
    gl64 = F.regress_bwd(fid, lg.double(), gp64)
    gx = np.empty((B, H, W, 3), np.float32)
    gl = np.zeros((B, PS), np.float32)
    assert hm.hm_bwd(fid, _p(xs), _p(gys), _p(gx), _p(gl), _p(lgs), PS, B, H * W, 1) == 0
    n = F.NUM_PARAMS[fid]
    eg = (torch.from_numpy(gl[:, :n]).double() - gl64).abs()
    scale = gl64.abs().max(dim=1, keepdim=True).values.clamp_min(1e-6)
    assert (eg <= 2e-4 * scale).all(), (fid, float((eg / scale).max()))
    # and the ranges really changed the result
    monkeypatch.undo()
    assert not torch.allclose(F.regress(fid, lg), p32)
  finally:
    hm.hm_set_ranges(0.0, 0.0, 0.0, 0.0, 0.0, 0.0)


def test_curve_knot_ties_follow_tf_clip(hm):
  """x exactly on knots / outside [0,1]: both neighbouring segments pass on a knot, none outside."""
  vals = [-0.5, -0.0, 0.0, 0.125, 0.25, 0.5, 0.875, 1.0, 1.0000001, 1.5, 0.3, 0.999]
  x = torch.tensor(vals, dtype=torch.float32).repeat_interleave(3).reshape(1, 1, len(vals), 3).contiguous()
  for fid in (F.T, F.C):
    lg = F.synth_logits(fid, 1, seed=99)
    gy = torch.ones_like(x)
    gx64, _ = F.process_bwd_autograd(fid, x.double(), F.regress(fid, lg.double()), gy.double())
    gx = np.empty(tuple(x.shape), np.float32)
    gl = np.zeros((1, PS), np.float32)
    xs, gys, lgs = _np(x), _np(gy), _pad(lg)
    assert hm.hm_bwd(fid, _p(xs), _p(gys), _p(gx), _p(gl), _p(lgs), PS, 1, len(vals), 1) == 0
    assert np.allclose(gx, gx64.numpy(), rtol=1e-5, atol=1e-6), (fid, gx.ravel()[::3], gx64.numpy().ravel()[::3])


@pytest.mark.parametrize("masking", [1, 0])
@pytest.mark.parametrize("shape", [(2, 16, 16), (2, 12, 20), (1, 21, 9)])
@pytest.mark.parametrize("fid", [F.E, F.SP, F.T, F.CT, F.C, F.LE, F.VG])
def test_masked_apply_math(hm, fid, shape, masking):
  B, H, W = shape
  x = F.synth_images(B, H, W, seed=60 + fid)
  lg = F.synth_logits(fid, B) * 0.7
  nm = 5 if fid == F.VG else 6
  ml = torch.randn(B, nm, generator=torch.Generator().manual_seed(3 + fid)) * 0.8
  gy = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(9 + fid))
  kw = dict(maximum_sharpness=1.5, minimum_strength=0.3)
  ref32 = F.apply_masked(fid, x, lg, ml, bool(masking), **kw)
  mask32 = F.get_mask(fid, x, ml, bool(masking), **kw)
  gx64, gl64, gm64 = F.apply_masked_bwd_autograd(fid, x.double(), lg.double(), ml.double(), gy.double(), bool(masking), **kw)
  xs, gys, lgs = _np(x), _np(gy), _pad(lg)
  mls = np.zeros((B, 6), np.float32)
  mls[:, :nm] = _np(ml)
  y = np.empty((B, H, W, 3), np.float32)
  mo = np.empty((B, H, W), np.float32)
  assert hm.hm_masked_fwd(fid, _p(xs), _p(y), _p(mo), _p(lgs), PS, _p(mls), 6, B, H, W,
                          ctypes.c_float(1.5), ctypes.c_float(0.3), masking, 1) == 0
  m_ref = mask32.expand(B, H, W, 1)[..., 0]
  assert np.allclose(mo, m_ref.numpy(), rtol=0, atol=3e-6)
  err = (torch.from_numpy(y) - ref32).abs()
  tol = 1e-5 * ref32.abs().clamp_min(1e-4) + 4e-6 * (F.process(fid, x, F.regress(fid, lg)) - x).abs()
  if fid == F.SP:                                    # closed-form ramps vs the fp32 hue round trip (see _fwd_tol)
    tol = tol + 1e-6 * x.clamp(max=1.0).amax(dim=-1, keepdim=True).abs()
  assert (err <= tol).all(), float((err / tol).max())
  gx = np.empty((B, H, W, 3), np.float32)
  gl = np.zeros((B, PS), np.float32)
  gm = np.zeros((B, 6), np.float32)
  assert hm.hm_masked_bwd(fid, _p(xs), _p(gys), _p(gx), _p(gl), _p(gm), _p(lgs), PS, _p(mls), 6, B, H, W,
                          ctypes.c_float(1.5), ctypes.c_float(0.3), masking, 1) == 0
  if fid != F.SP:     # TF defines no S+ image gradient; the closed form is checked in test_backward_math
    e = (torch.from_numpy(gx).double() - gx64).abs()
    assert (e <= 1e-4 * gx64.abs().clamp_min(1e-3 * float(gx64.abs().max()) + 1e-30)).all(), float(e.max())
  n = F.NUM_PARAMS[fid]
  for got, ref in ((gl[:, :n], gl64), (gm[:, :nm], gm64)):
    eg = (torch.from_numpy(got).double() - ref).abs()
    scale = ref.abs().max(dim=1, keepdim=True).values.clamp_min(1e-6)
    assert (eg <= 2e-4 * scale).all(), (fid, float((eg / scale).max()))


# ---- the same host-compiled device math against vectors made by the REFERENCE'S OWN CODE ---------------
# (tests/golden/reference_golden.npz sections 1-2, see tests/test_reference_golden.py): the CUDA source's
# per-pixel functions, regressors, finalisers and mask are held to filters.py directly, without a GPU and
# without the oracle in between.  fp32 device math vs the fixture's fp64 run: pixels 1e-5 (north_star),
# gradients 1e-4 of scale.
@pytest.fixture(scope="module")
def refgold():
  return np.load(os.path.join(HERE, "golden", "reference_golden.npz"))


def _rel(a, b, floor):
  a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
  s = np.abs(b).max()
  if s == 0:
    return float(np.abs(a).max())
  return float((np.abs(a - b) / np.maximum(np.abs(b), floor * s)).max())


@pytest.mark.parametrize("fid", range(10))
def test_device_math_matches_reference_code(hm, refgold, fid):
  p = "f%d_" % fid
  x, lg, gy = refgold[p + "x"], refgold[p + "logits"], refgold[p + "gy"]
  B, H, W, _ = x.shape
  n = lg.shape[1]
  xs = np.ascontiguousarray(x, np.float32)
  gys = np.ascontiguousarray(gy, np.float32)
  lgs = np.zeros((B, PS), np.float32)
  lgs[:, :n] = lg
  y = np.empty((B, H, W, 3), np.float32)
  assert hm.hm_fwd(fid, _p(xs), _p(y), _p(lgs), PS, B, H * W, 1) == 0
  want = refgold[p + "y"]
  if fid == F.CT:
    xt = torch.from_numpy(x)
    lum = F.rgb2lum(xt).clamp(0, 1).numpy()
    prm = np.abs(refgold[p + "param"]).reshape(-1, 1, 1, 1)
    tol = 1e-5 * np.maximum(np.abs(want), 1e-4) + prm * np.abs(x) / (lum + 1e-6) * (2 * 2.0 ** -24) + 1e-7 * np.abs(want).max()
    assert (np.abs(y - want) <= tol).all()
  else:
    assert _rel(y, want, 1e-2) < 1e-5
  gx = np.empty((B, H, W, 3), np.float32)
  gl = np.zeros((B, PS), np.float32)
  assert hm.hm_bwd(fid, _p(xs), _p(gys), _p(gx), _p(gl), _p(lgs), PS, B, H * W, 1) == 0
  want_gx = refgold[p + "gx"]
  if fid == F.SP:          # away from exact channel ties (TF 1.6 has no RGBToHSV gradient; tie conventions differ)
    px = np.minimum(x, 1.0).reshape(-1, 3)
    keep = (px[:, 0] != px[:, 1]) & (px[:, 1] != px[:, 2]) & (px[:, 0] != px[:, 2])
    assert _rel(gx.reshape(-1, 3)[keep], want_gx.reshape(-1, 3)[keep], 1e-3) < 1e-4
  else:
    assert _rel(gx, want_gx, 1e-3) < 1e-4
  assert _rel(gl[:, :n], refgold[p + "glogits"], 1e-2) < 2e-4


@pytest.mark.parametrize("fid", range(10))
def test_masked_device_math_matches_reference_code(hm, refgold, fid):
  p = "m%d_" % fid
  n = F.NUM_PARAMS[fid]
  x, feat, gy = refgold[p + "x"], refgold[p + "feat"], refgold[p + "gy"]
  W1, b1, W2, b2 = (refgold[p + k] for k in ("fc1_weights", "fc1_biases", "fc2_weights", "fc2_biases"))
  h = feat @ W1 + b1
  h = 0.6 * h + 0.4 * np.abs(h)
  o = h @ W2 + b2                                    # extract_parameters on the host (fp64)
  B, H, W, _ = x.shape
  nm = o.shape[1] - n
  xs, gys = np.ascontiguousarray(x, np.float32), np.ascontiguousarray(gy, np.float32)
  lgs = np.zeros((B, PS), np.float32)
  lgs[:, :n] = o[:, :n]
  mls = np.zeros((B, 6), np.float32)
  mls[:, :nm] = o[:, n:]
  y = np.empty((B, H, W, 3), np.float32)
  mo = np.empty((B, H, W), np.float32)
  assert hm.hm_masked_fwd(fid, _p(xs), _p(y), _p(mo), _p(lgs), PS, _p(mls), 6, B, H, W,
                          ctypes.c_float(float(1)), ctypes.c_float(0.3), 1, 1) == 0        # cfg.maximum_sharpness = 1
  assert _rel(mo, refgold[p + "mask"][..., 0], 1e-3) < 1e-5
  assert _rel(y, refgold[p + "low"], 1e-2) < 2e-5
  if fid == F.SP:
    return
  gx = np.empty((B, H, W, 3), np.float32)
  gl = np.zeros((B, PS), np.float32)
  gm = np.zeros((B, 6), np.float32)
  assert hm.hm_masked_bwd(fid, _p(xs), _p(gys), _p(gx), _p(gl), _p(gm), _p(lgs), PS, _p(mls), 6, B, H, W,
                          ctypes.c_float(float(1)), ctypes.c_float(0.3), 1, 1) == 0
  assert _rel(gx, refgold[p + "gx"], 1e-3) < 1e-4
  go = np.concatenate([gl[:, :n], gm[:, :nm]], axis=1).astype(np.float64)
  assert _rel(go.sum(axis=0), refgold[p + "g_fc2_biases"], 1e-2) < 3e-4
  assert _rel(h.T @ go, refgold[p + "g_fc2_weights"], 1e-2) < 3e-4
