"""Inference / high-resolution path (net.py:683-877, evaluate.py:8-31).

The reference resizes the centre crop of each linear image to 64x64, runs cfg.test_steps policy
steps (one sess.run each, batch size 1) and applies every step's selected filter to the
full-resolution image inside the same sess.run.  Here the policy episode runs on the batch of
thumbnails first (is_train = 0 -> argmax action, dropout still active like the reference,
agent.py:36), recording per step the selected filter id and its raw regressor logits; then ALL
steps are applied to the full-resolution batch by ONE fused kernel (exp_filter_chain_fwd: 24
B/pixel for the whole episode).  Images of equal resolution are batched (evaluate.py:14-18 lists
batching as a TODO)."""
import torch
import torch.nn.functional as Fn

from . import ops


def center_thumbnail(high_res, size=64):
  """util.get_image_center + cv2.resize(..., (64, 64)) (net.py:779): centre square crop, bilinear
  (half-pixel centres, no anti-aliasing -- cv2.INTER_LINEAR semantics)."""
  B, H, W, _ = high_res.shape
  s = min(H, W)
  y0, x0 = (H - s) // 2, (W - s) // 2
  crop = high_res[:, y0:y0 + s, x0:x0 + s, :].permute(0, 3, 1, 2)
  thumb = Fn.interpolate(crop, size=(size, size), mode="bilinear", align_corners=False, antialias=False)
  return thumb.permute(0, 2, 3, 1).contiguous()


def retouch(trainer, high_res, generator=None, steps=None, fused=True):
  """high_res: [B,H,W,3] linear RGB on the GPU.  Returns dict(output, ids [S,B], logits [S,B,24],
  thumbnails, states).  `fused=False` applies the steps one kernel at a time (reference order of
  operations; used by the tests to check the fused kernel)."""
  cfg = trainer.cfg
  B = high_res.shape[0]
  dev = high_res.device
  S = steps or cfg.test_steps
  thumb = center_thumbnail(high_res, cfg.source_img_size)
  states = torch.zeros(B, cfg.num_state_dim, device=dev)
  masking = bool(getattr(cfg, "masking", False))
  ids, logits, mask_logits = [], [], []
  for _ in range(S):
    noise, drop_f, drop_s, _ = trainer.draw(B, generator)
    c = trainer.policy.forward(thumb, states, noise, drop_f, drop_s, 0, 0.0, cfg)
    ids.append(c.ids)
    logits.append(c.logits_sel)
    if masking:
      mask_logits.append(c.mask_logits_sel)
    thumb, states = c.out, c.new_states
  ids = torch.stack(ids).contiguous()
  logits = torch.stack(logits).contiguous()
  if masking:
    # cfg.masking: each step's mask depends on the full-resolution pixels' luminance AFTER the
    # previous steps (filters.py:92-96), so the steps run as S masked launches
    out = high_res.contiguous()
    for s in range(S):
      out = ops.filter_masked_fwd(out, logits[s], mask_logits[s], ids[s], float(cfg.maximum_sharpness),
                                  float(cfg.minimum_strength), True, out=torch.zeros_like(out), logits=True)
  elif fused:
    out = ops.filter_chain_fwd(high_res.contiguous(), logits, ids, logits=True)
  else:
    out = high_res.contiguous()
    for s in range(S):
      p = ops.filter_regress_fwd(logits[s], ids[s])
      out = ops.filter_fwd(out, p, ids[s])
  return dict(output=out, ids=ids, logits=logits, thumbnails=thumb, states=states)


# ---- image I/O around the hot path (net.py:731-747, 825-877; util.py:311-323, 495-501) --------
def load_linear_image(path):
  """Reads a tif/tiff (ProPhoto RGB, linearised with x^1.8 like util.linearize_ProPhotoRGB) or a
  jpg/png (sRGB: x^2.2 then / (2 max), net.py:739-747) into a float32 HxWx3 RGB array."""
  import cv2
  import numpy as np
  img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
  if img is None:
    raise IOError("cannot read %s" % path)
  if img.ndim == 2:
    img = np.stack([img] * 3, axis=2)
  img = img[:, :, :3][:, :, ::-1]
  depth = 16 if img.dtype == np.uint16 else 8
  img = img.astype(np.float32) * (1.0 / (2 ** depth - 1))          # util.read_tiff16
  if path.lower().endswith((".tif", ".tiff")):
    return np.power(img, 1.8).astype(np.float32)
  lin = np.power(img, 2.2)
  return (lin / (2 * lin.max())).astype(np.float32)


def save_png(path, img):
  """net.py:771-774 show_and_save: RGB float [0,1] -> 8-bit PNG."""
  import cv2
  import numpy as np
  cv2.imwrite(path, np.clip(img[:, :, ::-1] * 255.0, 0, 255).astype(np.uint8))


def evaluate_files(trainer, files, output_dir="./outputs", generator=None):
  """GAN.eval (net.py:711-877) on a list of image files: images of equal resolution are retouched
  as one batch.  Writes <name>.retouched.png; returns {file: (ids per step)}."""
  import os
  import numpy as np
  os.makedirs(output_dir, exist_ok=True)
  groups = {}
  for fn in files:
    im = load_linear_image(fn)
    groups.setdefault(im.shape[:2], []).append((fn, im))
  result = {}
  for res, items in groups.items():
    batch = torch.from_numpy(np.stack([im for _, im in items])).to(trainer.device)
    out = retouch(trainer, batch, generator=generator)
    o = out["output"].cpu().numpy()
    for k, (fn, _) in enumerate(items):
      save_png(os.path.join(output_dir, os.path.basename(fn) + ".retouched.png"), o[k])
      result[fn] = out["ids"][:, k].tolist()
  return result
