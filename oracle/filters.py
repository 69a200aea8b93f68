"""CPU oracle for the per-pixel filter stack -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, op for op, the arithmetic of the reference's ``filters.py`` on torch CPU
tensors.  The same code runs in float32 (the reference's execution dtype; this is the
"CPU restatement of the reference TF graph") and in float64 (the arbiter used when the
fp32 formula itself is ill-conditioned, e.g. ContrastFilter's ``-cos(pi*l)*0.5+0.5``).

PARITY: pinned to the reference's own filters.py executed over a TF-1 API stand-in
(tests/test_reference_golden.py; see oracle/__init__.py), not to TF binaries.
Numerics that live in TensorFlow 1.6 rather than in the repo
(``tf.image.rgb_to_hsv`` / ``hsv_to_rgb`` -- tensorflow/core/kernels/colorspace_op.h,
``tf.clip_by_value`` tie rules, ``tf.maximum``/``tf.minimum`` gradients) are restated
from TF's published kernel definitions.

Layout: images are NHWC ``[B, H, W, 3]``; ``param`` is the *regressed* filter parameter
(post ``filter_param_regressor``) flattened to ``[B, n]``:

    id  class                        n   param
    0   ExposureFilter               1   p  (EV)                    filters.py:170-182
    1   GammaFilter                  1   gamma                      filters.py:194-206
    2   ImprovedWhiteBalanceFilter   3   s_r, s_g, s_b              filters.py:215-238
    3   SaturationPlusFilter         1   p                          filters.py:474-498
    4   ToneFilter                   8   t_0..t_7                   filters.py:298-322
    5   ContrastFilter               1   p                          filters.py:404-419
    6   WNBFilter                    1   p                          filters.py:428-440
    7   ColorFilter                 24   t_{c,i} at index c*8+i     filters.py:247-273
    8   LevelFilter                  2   lower, upper-1             filters.py:449-464
    9   VignetFilter                 1   (unused: process == img*0) filters.py:341-352

Ids 0..7 are ``cfg.filters`` in config_example.py:22-25; 8 and 9 are the two Filter subclasses
no shipped config lists.  ``get_mask`` / ``apply_masked`` restate Filter.apply with
cfg.masking == True (filters.py:62-99, 110-148; VignetFilter.get_mask filters.py:354-396).
"""
import math

import numpy as np
import torch

# --------------------------------------------------------------------------------------
# constants (config_example.py:27-33) and helpers (util.py)
# --------------------------------------------------------------------------------------
FILTER_NAMES = ["E", "G", "W", "S+", "T", "Ct", "BW", "C", "Le", "V"]
NUM_PARAMS = [1, 1, 3, 1, 8, 1, 1, 24, 2, 1]
E, G, W, SP, T, CT, BW, C, LE, VG = range(10)
MASK_PARAMS = 6                 # Filter.get_num_mask_parameters (filters.py:107-108); Vignet uses 5
MAXIMUM_SHARPNESS = 1           # cfg.maximum_sharpness (config_example.py:38)
MINIMUM_STRENGTH = 0.3          # cfg.minimum_strength  (config_example.py:37)

CURVE_STEPS = 8                 # cfg.curve_steps
EXPOSURE_RANGE = 3.5            # cfg.exposure_range
GAMMA_RANGE = 3                 # cfg.gamma_range
TONE_CURVE_RANGE = (0.5, 2)     # cfg.tone_curve_range
COLOR_CURVE_RANGE = (0.90, 1.10)  # cfg.color_curve_range


def f32c(v):
  """A python float as TF would embed it in a float32 graph (rounded to fp32)."""
  return float(np.float32(v))


LN2 = f32c(np.log(2))           # filters.py:182  ``* np.log(2)``
PI = f32c(math.pi)              # filters.py:417  ``math.pi * luminance``
LUM_R, LUM_G, LUM_B = f32c(0.27), f32c(0.67), f32c(0.06)   # util.py:271-274


def tf_maximum(x, c):
  """tf.maximum(x, const): forward max; gradient goes to x where x >= c (TF MaximumGrad)."""
  return torch.where(x >= c, x, torch.full_like(x, c))


def tf_minimum(x, c):
  """tf.minimum(x, const): gradient goes to x where x <= c (TF MinimumGrad)."""
  return torch.where(x <= c, x, torch.full_like(x, c))


def tf_clip(x, lo, hi):
  """tf.clip_by_value in TF 1.6 == minimum(maximum(x, lo), hi)."""
  return tf_minimum(tf_maximum(x, lo), hi)


def rgb2lum(img):
  """util.py:271-274."""
  lum = LUM_R * img[..., 0] + LUM_G * img[..., 1] + LUM_B * img[..., 2]
  return lum[..., None]


def lerp(a, b, l):
  """util.py:307-308."""
  return (1 - l) * a + l * b


def tanh_range(l, r, initial=None):
  """util.py:281-294."""
  if initial is not None:
    bias = math.atanh(2 * (initial - l) / (r - l) - 1)
  else:
    bias = 0

  def activation(x):
    return (torch.tanh(x + bias) * 0.5 + 0.5) * (r - l) + l

  return activation


# --------------------------------------------------------------------------------------
# filter_param_regressor: logits [B, n] -> param [B, n]
# --------------------------------------------------------------------------------------
def regress(fid, f):
  if fid == E:     # filters.py:177-179
    return tanh_range(-EXPOSURE_RANGE, EXPOSURE_RANGE, initial=0)(f)
  if fid == G:     # filters.py:201-203
    lg = float(np.log(GAMMA_RANGE))
    return torch.exp(tanh_range(-lg, lg)(f))
  if fid == W:     # filters.py:223-235
    mask = torch.tensor([[0.0, 1.0, 1.0]], dtype=f.dtype)
    f = f * mask
    s = torch.exp(tanh_range(-0.5, 0.5)(f))
    s = s * (1.0 / (1e-5 + LUM_R * s[:, 0] + LUM_G * s[:, 1] + LUM_B * s[:, 2]))[:, None]
    return s
  if fid in (SP, BW, LE, VG):   # filters.py:481-482, 435-436, 456-457, 348-349
    return torch.sigmoid(f)
  if fid == T:     # filters.py:306-310
    return tanh_range(*TONE_CURVE_RANGE)(f)
  if fid == CT:    # filters.py:411-413
    return torch.tanh(f)
  if fid == C:     # filters.py:256-262 (features reshaped (-1, channels, curve_steps))
    return tanh_range(*COLOR_CURVE_RANGE, initial=1)(f)
  raise ValueError(fid)


# --------------------------------------------------------------------------------------
# TF colour-space kernels (tensorflow/core/kernels/colorspace_op.h), restated
# --------------------------------------------------------------------------------------
def _safe(d):
  return torch.where(d != 0, d, torch.ones_like(d))


def rgb_to_hsv(img):
  r, g, b = img[..., 0], img[..., 1], img[..., 2]
  # max / min with an explicit first-index tie rule so that autograd is well defined
  is_r = (r >= g) & (r >= b)
  is_g = (~is_r) & (g >= b)
  v = torch.where(is_r, r, torch.where(is_g, g, b))
  mn_r = (r <= g) & (r <= b)
  mn_g = (~mn_r) & (g <= b)
  mn = torch.where(mn_r, r, torch.where(mn_g, g, b))
  rng = v - mn
  s = torch.where(v > 0, rng / _safe(v), torch.zeros_like(v))
  norm = (1.0 / _safe(rng)) * f32c(1.0 / 6.0)
  h = torch.where(r == v, norm * (g - b),
                  torch.where(g == v, norm * (b - r) + f32c(2.0 / 6.0),
                              norm * (r - g) + f32c(4.0 / 6.0)))
  h = torch.where(rng > 0, h, torch.zeros_like(h))
  h = torch.where(h < 0, h + 1, h)
  return h, s, v


def hsv_to_rgb(h, s, v):
  dh = h * 6
  dr = tf_clip(torch.abs(dh - 3) - 1, 0.0, 1.0)
  dg = tf_clip(-torch.abs(dh - 2) + 2, 0.0, 1.0)
  db = tf_clip(-torch.abs(dh - 4) + 2, 0.0, 1.0)
  one_s = -s + 1
  return torch.stack([(one_s + s * dr) * v, (one_s + s * dg) * v, (one_s + s * db) * v], dim=-1)


# --------------------------------------------------------------------------------------
# process(img, param): the per-pixel op of each filter, reference op order
# --------------------------------------------------------------------------------------
def _curve(img, t):
  """filters.py:264-273 / 312-322.  t: [B, 1 or 3, 8] broadcast over channels."""
  L = CURVE_STEPS
  tt = t[:, None, None, :, :]
  curve_sum = tt.sum(dim=4) + 1e-30
  total = img * 0
  for i in range(L):
    total = total + tf_clip(img - 1.0 * i / L, 0.0, 1.0 / L) * tt[..., i]
  total = total * (L / curve_sum)
  return total


def process(fid, img, param):
  B = img.shape[0]
  if fid == E:      # filters.py:181-182
    return img * torch.exp(param[:, None, None, :] * LN2)
  if fid == G:      # filters.py:205-206
    return torch.pow(tf_maximum(img, 0.001), param[:, None, None, :])
  if fid == W:      # filters.py:237-238
    return img * param[:, None, None, :]
  if fid == SP:     # filters.py:484-498
    xm = tf_minimum(img, 1.0)
    h, s, v = rgb_to_hsv(xm)
    enhanced_s = s + (1 - s) * (0.5 - torch.abs(0.5 - v)) * 0.8
    full = hsv_to_rgb(h, enhanced_s, v)
    p = param[:, :, None, None]
    return xm * (1.0 - p) + full * p
  if fid == T:      # filters.py:312-322
    return _curve(img, param.reshape(B, 1, CURVE_STEPS))
  if fid == CT:     # filters.py:415-419
    lum = tf_minimum(tf_maximum(rgb2lum(img), 0.0), 1.0)
    contrast_lum = -torch.cos(PI * lum) * 0.5 + 0.5
    contrast_image = img / (lum + 1e-6) * contrast_lum
    return lerp(img, contrast_image, param[:, :, None, None])
  if fid == BW:     # filters.py:438-440
    return lerp(img, rgb2lum(img), param[:, :, None, None])
  if fid == C:      # filters.py:264-273
    return _curve(img, param.reshape(B, 3, CURVE_STEPS))
  if fid == LE:     # filters.py:459-464
    lower = param[:, 0][:, None, None, None]
    upper = (param[:, 1] + 1)[:, None, None, None]
    return tf_clip((img - lower) / (upper - lower + 1e-6), 0.0, 1.0)
  if fid == VG:     # filters.py:351-352
    return img * 0
  raise ValueError(fid)


def mask_grid(H, W):
  """The centred unit grid of filters.py:126-135 (computed in float64 by numpy, stored float32)."""
  grid = np.zeros(shape=[1, H, W, 2], dtype=np.float32)
  shorter_edge = min(H, W)
  ii = (np.arange(H) + (shorter_edge - H) / 2.0) / shorter_edge - 0.5
  jj = (np.arange(W) + (shorter_edge - W) / 2.0) / shorter_edge - 0.5
  grid[0, :, :, 0] = ii[:, None]
  grid[0, :, :, 1] = jj[None, :]
  return grid


def get_mask(fid, img, mask_logits, masking=True, maximum_sharpness=MAXIMUM_SHARPNESS,
             minimum_strength=MINIMUM_STRENGTH):
  """Filter.get_mask (filters.py:110-148) or, for fid == VG, VignetFilter.get_mask
  (filters.py:354-396).  mask_logits [B, 6 | 5] raw; returns [B,H,W,1] (or ones(1,1,1,1))."""
  if fid != VG and not masking:
    return torch.ones(1, 1, 1, 1, dtype=img.dtype)
  filter_input_range = 5
  mp = tanh_range(l=-filter_input_range, r=filter_input_range, initial=0)(mask_logits)
  H, Wd = img.shape[1:3]
  grid = torch.from_numpy(mask_grid(H, Wd)).to(img.dtype)
  if fid != VG:
    inp = grid[:, :, :, 0, None] * mp[:, None, None, 0, None] + \
          grid[:, :, :, 1, None] * mp[:, None, None, 1, None] + \
          mp[:, None, None, 2, None] * (rgb2lum(img) - 0.5) + \
          mp[:, None, None, 3, None] * 2
    inp = inp * (maximum_sharpness * mp[:, None, None, 4, None] / filter_input_range)
    mask = torch.sigmoid(inp)
    mask = mask * (mp[:, None, None, 5, None] / filter_input_range * 0.5 + 0.5) * (1 - minimum_strength) + minimum_strength
    return mask
  inp = (grid[:, :, :, 0, None] * mp[:, None, None, 0, None]) ** 2 + \
        (grid[:, :, :, 1, None] * mp[:, None, None, 1, None]) ** 2 + \
        mp[:, None, None, 2, None] - filter_input_range
  inp = inp * (maximum_sharpness * mp[:, None, None, 3, None] / filter_input_range)
  mask = torch.sigmoid(inp)
  mask = mask * (mp[:, None, None, 4, None] / filter_input_range * 0.5 + 0.5)
  if not masking:
    mask = mask * 0 + 1
  return mask


def apply_masked(fid, img, logits, mask_logits, masking=True, **kw):
  """Filter.apply (filters.py:62-99): lerp(img, process(img, regress(logits)), get_mask(...))."""
  param = regress(fid, logits)
  return lerp(img, process(fid, img, param), get_mask(fid, img, mask_logits, masking, **kw))


def apply_masked_bwd_autograd(fid, img, logits, mask_logits, gy, masking=True, **kw):
  """(gx, glogits, gmask_logits) of <gy, apply_masked(...)> by autograd of the restatement."""
  x = img.detach().clone().requires_grad_(True)
  f = logits.detach().clone().requires_grad_(True)
  m = mask_logits.detach().clone().requires_grad_(True)
  y = apply_masked(fid, x, f, m, masking, **kw)
  gx, gf, gm = torch.autograd.grad(y, [x, f, m], grad_outputs=gy, allow_unused=True)
  z = lambda g, t: torch.zeros_like(t) if g is None else g
  return z(gx, x), z(gf, f), z(gm, m)


def apply_filter(fid, img, logits):
  """Filter.apply with masking disabled (filters.py:62-99; mask == ones(1,1,1,1))."""
  param = regress(fid, logits)
  return lerp(img, process(fid, img, param), torch.ones(1, 1, 1, 1, dtype=img.dtype))


# --------------------------------------------------------------------------------------
# backward: autograd of the restatement (reference gradient definition) ...
# --------------------------------------------------------------------------------------
def process_bwd_autograd(fid, img, param, gy):
  """d<gy, process(img,param)>/d(img, param) via torch autograd of the restatement.

  This is what tf.gradients would build for every filter except S+ w.r.t. the image
  (TF 1.6 registers no gradient for RGBToHSV / HSVToRGB)."""
  x = img.detach().clone().requires_grad_(True)
  p = param.detach().clone().requires_grad_(True)
  y = process(fid, x, p)
  gx, gp = torch.autograd.grad(y, [x, p], grad_outputs=gy)
  return gx, gp


# ... and the hand-derived closed forms the CUDA kernels implement
def process_bwd_analytic(fid, img, param, gy):
  """Closed-form d<gy, y>/d(img, param).  Same dtype as the inputs.  Returns (gx, gparam)."""
  B = img.shape[0]
  x = img
  L = CURVE_STEPS
  sum_hw = lambda t: t.sum(dim=(1, 2))
  if fid == E:
    e = torch.exp(param * LN2)[:, None, None, :]
    y = x * e
    return gy * e, sum_hw((gy * y).sum(-1, keepdim=True)) * LN2
  if fid == G:
    g = param[:, None, None, :]
    xc = torch.clamp_min(x, 0.001)
    y = torch.pow(xc, g)
    gx = gy * g * y / xc * (x >= 0.001).to(x.dtype)
    return gx, sum_hw((gy * y * torch.log(xc)).sum(-1, keepdim=True))
  if fid == W:
    return gy * param[:, None, None, :], sum_hw(gy * x)
  if fid == SP:
    p = param[:, :, None, None]
    xm = torch.clamp_max(x, 1.0)
    r, g, b = xm[..., 0], xm[..., 1], xm[..., 2]
    is_r = (r >= g) & (r >= b)
    is_g = (~is_r) & (g >= b)
    is_b = ~(is_r | is_g)
    mn_r = (r <= g) & (r <= b)
    mn_g = (~mn_r) & (g <= b)
    mn_b = ~(mn_r | mn_g)
    V = torch.where(is_r, r, torch.where(is_g, g, b))[..., None]
    m = torch.where(mn_r, r, torch.where(mn_g, g, b))[..., None]
    rng = V - m
    k = (0.5 - torch.abs(0.5 - V)) * 0.8
    kp = 0.8 * torch.sign(0.5 - V)
    full = process(SP, x, torch.ones_like(param))        # p == 1 -> full colour image
    gparam = sum_hw((gy * (full - xm)).sum(-1, keepdim=True))
    gF = gy * p
    deg = (rng <= 0)
    rs = torch.where(deg, torch.ones_like(rng), rng)
    u = (V - xm) / rs
    Q = k * m
    Tt = (gF * u).sum(-1, keepdim=True)
    sgF = gF.sum(-1, keepdim=True)
    # generic pixel: F_c = xm_c - Q u_c
    gxm = gy + Q * gF / rs
    gV = -Tt * kp * m - Q * (sgF - Tt) / rs
    gm = -Tt * k - Q * Tt / rs
    # degenerate (grey) pixel: F = (V, (1-k)V, (1-k)V) as TF's hue==0 produces
    gxm_deg = gy * (1 - p)
    gV_deg = gF[..., 0:1] + (gF[..., 1:2] + gF[..., 2:3]) * (1 - k - kp * V)
    gxm = torch.where(deg, gxm_deg, gxm)
    gV = torch.where(deg, gV_deg, gV)
    gm = torch.where(deg, torch.zeros_like(gm), gm)
    amax = torch.stack([is_r, is_g, is_b], dim=-1).to(x.dtype)
    amin = torch.stack([mn_r, mn_g, mn_b], dim=-1).to(x.dtype)
    gxm = gxm + gV * amax + gm * amin
    return gxm * (x <= 1.0).to(x.dtype), gparam
  if fid in (T, C):
    nch = 1 if fid == T else 3
    t = param.reshape(B, nch, L)[:, None, None, :, :]            # [B,1,1,nch,8]
    S = t.sum(-1) + 1e-30                                        # [B,1,1,nch]
    y = process(fid, x, param)
    xe = x[..., None]                                            # [B,H,W,3,1]
    knots = torch.arange(L, dtype=x.dtype) / L
    v = xe - knots                                               # [B,H,W,3,8]
    clip = torch.clamp(v, 0.0, 1.0 / L)
    passes = ((v >= 0) & (v <= 1.0 / L)).to(x.dtype)
    slope = (passes * t).sum(-1) * (L / S)
    gx = gy * slope
    A = (gy[..., None] * clip)                                   # [B,H,W,3,8]
    Bs = (gy * y)
    if nch == 1:
      A = A.sum(dim=(1, 2, 3))                                   # [B,8]
      Bs = Bs.sum(dim=(1, 2, 3))[:, None]
      gparam = (L * A - Bs) / S.reshape(B, 1)
    else:
      A = A.sum(dim=(1, 2))                                      # [B,3,8]
      Bs = Bs.sum(dim=(1, 2))[:, :, None]
      gparam = ((L * A - Bs) / S.reshape(B, 3, 1)).reshape(B, 3 * L)
    return gx, gparam
  if fid == CT:
    p = param[:, :, None, None]
    lum_raw = rgb2lum(x)
    l = torch.clamp(lum_raw, 0.0, 1.0)
    inside = ((lum_raw >= 0) & (lum_raw <= 1)).to(x.dtype)
    cl = torch.sin(0.5 * PI * l) ** 2         # == -cos(pi l)/2 + 1/2, well conditioned
    dcl = 0.5 * PI * torch.sin(PI * l)
    den = l + 1e-6
    w = cl / den
    dw = dcl / den - cl / (den * den)
    coef = torch.tensor([LUM_R, LUM_G, LUM_B], dtype=x.dtype)
    sgx = (gy * x).sum(-1, keepdim=True)
    gx = gy * ((1 - p) + p * w) + p * sgx * dw * inside * coef
    gparam = sum_hw((gy * (x * w - x)).sum(-1, keepdim=True))
    return gx, gparam
  if fid == BW:
    p = param[:, :, None, None]
    lum = rgb2lum(x)
    coef = torch.tensor([LUM_R, LUM_G, LUM_B], dtype=x.dtype)
    sg = gy.sum(-1, keepdim=True)
    gx = (1 - p) * gy + p * sg * coef
    gparam = sum_hw((gy * (lum - x)).sum(-1, keepdim=True))
    return gx, gparam
  if fid == LE:
    lo = param[:, 0][:, None, None, None]
    d = (param[:, 1] + 1)[:, None, None, None] - lo + 1e-6
    v = (x - lo) / d
    g = gy * ((v >= 0) & (v <= 1)).to(x.dtype)              # TF clip_by_value passes ties
    gparam = torch.stack([(g * (v - 1) / d).sum(dim=(1, 2, 3)), (g * (-v) / d).sum(dim=(1, 2, 3))], dim=1)
    return g / d, gparam
  if fid == VG:
    return gy * 0, torch.zeros_like(param)
  raise ValueError(fid)


def regress_bwd(fid, logits, gparam):
  """d<gparam, regress(logits)>/dlogits by autograd (per-image, tiny)."""
  f = logits.detach().clone().requires_grad_(True)
  p = regress(fid, f)
  (gf,) = torch.autograd.grad(p, [f], grad_outputs=gparam)
  return gf


# --------------------------------------------------------------------------------------
# chain driver: N filter steps applied in sequence (BASELINE.json configs[1]/[4])
# --------------------------------------------------------------------------------------
def chain_fwd(ids, img, logits_list):
  """ids: sequence of filter ids; logits_list[k]: [B, n_k].  Returns list of activations
  x_0..x_N (x_0 = img)."""
  acts = [img]
  for fid, f in zip(ids, logits_list):
    acts.append(process(fid, acts[-1], regress(fid, f)))
  return acts


def chain_fwd_bwd(ids, img, logits_list, gout, analytic=True):
  """Forward + backward through the chain.  Returns (y, gimg, [glogits_k])."""
  acts = chain_fwd(ids, img, logits_list)
  g = gout
  glogits = [None] * len(ids)
  for k in reversed(range(len(ids))):
    fid = ids[k]
    param = regress(fid, logits_list[k])
    fn = process_bwd_analytic if analytic else process_bwd_autograd
    g, gparam = fn(fid, acts[k], param, g)
    glogits[k] = regress_bwd(fid, logits_list[k], gparam)
  return acts[-1], g, glogits


# --------------------------------------------------------------------------------------
# synthetic inputs: dark linear-RGB-like batches with clamp/knot stressors
# --------------------------------------------------------------------------------------
def synth_images(B, H, W, seed=1234, dtype=torch.float32, stress=True):
  g = torch.Generator().manual_seed(seed)
  x = torch.exp(torch.randn(B, H, W, 3, generator=g, dtype=torch.float32) * 1.0 - 3.2).clamp_(0, 4)
  if stress:
    u = torch.rand(B, H, W, 3, generator=g)
    hi = 1 + 3 * torch.rand(B, H, W, 3, generator=g)
    x = torch.where(u < 0.010, hi, x)
    x = torch.where((u >= 0.010) & (u < 0.011), torch.zeros_like(x), x)
    x = torch.where((u >= 0.011) & (u < 0.012), torch.full_like(x, 0.001), x)
    knot = torch.randint(0, 9, (B, H, W, 3), generator=g).float() / 8
    x = torch.where((u >= 0.012) & (u < 0.013), knot, x)
  return x.to(dtype)


def synth_logits(fid, B, seed=4321, dtype=torch.float32):
  g = torch.Generator().manual_seed(seed + 17 * fid)
  return torch.randn(B, NUM_PARAMS[fid], generator=g, dtype=torch.float32).to(dtype)
