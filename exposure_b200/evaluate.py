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


def retouch(trainer, high_res, generator=None, steps=None, fused=True, trace=False):
  """high_res: [B,H,W,3] linear RGB on the GPU.  Returns dict(output, ids [S,B], logits [S,B,24],
  thumbnails, states).  `fused=False` applies the steps one kernel at a time (reference order of
  operations; used by the tests to check the fused kernel).  trace=True additionally records what the
  reference's eval loop keeps per step (net.py:785-820): `trajectory` (the S+1 thumbnails), `debug`
  ([S][B] host debug_info dicts) and `intermediate` (the S full-resolution images, step by step)."""
  cfg = trainer.cfg
  B = high_res.shape[0]
  dev = high_res.device
  # net.py:796-820 loops cfg.test_steps times and breaks at the first step whose new state is STOPPED; every image
  # of a batch stops at the same step (step + 1 == cfg.test_steps, agent.py:210-218), so more steps are never run
  S = min(steps or cfg.test_steps, cfg.test_steps)
  thumb = center_thumbnail(high_res, cfg.source_img_size)
  states = torch.zeros(B, cfg.num_state_dim, device=dev)
  masking = bool(getattr(cfg, "masking", False))
  ids, logits, mask_logits = [], [], []
  trajectory, debug = [thumb], []
  if trace:
    from .agent import host_debug_info
    filters = [cls(thumb, cfg) for cls in cfg.filters]
  for _ in range(S):
    noise, drop_f, drop_s, _ = trainer.draw(B, generator)
    c = trainer.policy.forward(thumb, states, noise, drop_f, drop_s, 0, 0.0, cfg)
    ids.append(c.ids)
    logits.append(c.logits_sel)
    if masking:
      mask_logits.append(c.mask_logits_sel)
    if trace:
      debug.append([host_debug_info(c, thumb, states, filters, cfg, k) for k in range(B)])
    thumb, states = c.out, c.new_states
    trajectory.append(thumb)
  ids = torch.stack(ids).contiguous()
  logits = torch.stack(logits).contiguous()
  intermediate = []
  if masking:
    # cfg.masking: each step's mask depends on the full-resolution pixels' luminance AFTER the
    # previous steps (filters.py:92-96), so the steps run as S masked launches
    out = high_res.contiguous()
    for s in range(S):
      out = ops.filter_masked_fwd(out, logits[s], mask_logits[s], ids[s], float(cfg.maximum_sharpness),
                                  float(cfg.minimum_strength), True, logits=True)
      intermediate.append(out)
  elif fused and not trace:
    out = ops.filter_chain_fwd(high_res.contiguous(), logits, ids, logits=True)
  else:
    out = high_res.contiguous()
    for s in range(S):
      p = ops.filter_regress_fwd(logits[s], ids[s])
      out = ops.filter_fwd(out, p, ids[s])
      intermediate.append(out)
  res = dict(output=out, ids=ids, logits=logits, thumbnails=thumb, states=states)
  if trace:
    res.update(trajectory=trajectory, debug=debug, intermediate=intermediate)
  return res


# ---- image I/O around the hot path (net.py:731-747, 825-877; util.py:311-323, 495-501) --------
def load_linear_image(path):
  """Reads a tif/tiff (ProPhoto RGB, linearised with x^1.8 like util.linearize_ProPhotoRGB) or a
  jpg/png (sRGB: x^2.2 then / (2 max), net.py:739-747) into a float32 HxWx3 RGB array."""
  import cv2
  import numpy as np
  # net.py:731-747: tiffs through util.read_tiff16 (full bit depth), every other format through the default
  # cv2.imread(fn) = 8-bit 3-channel BGR (a 16-bit png is reduced to 8 bit first, like the reference)
  is_tiff = path.lower().endswith((".tif", ".tiff"))
  img = cv2.imread(path, cv2.IMREAD_UNCHANGED) if is_tiff else cv2.imread(path)
  if img is None:
    raise IOError("cannot read %s" % path)
  if img.ndim == 2:
    img = np.stack([img] * 3, axis=2)
  img = img[:, :, :3][:, :, ::-1]
  depth = 16 if img.dtype == np.uint16 else 8
  img = img.astype(np.float32) * (1.0 / (2 ** depth - 1))          # util.read_tiff16
  if is_tiff:
    return np.power(img, 1.8).astype(np.float32)
  lin = np.power(img, 2.2)
  return (lin / (2 * lin.max())).astype(np.float32)


def save_png(path, img):
  """net.py:771-774 show_and_save: `cv2.imwrite(path, img[:, :, ::-1] * 255.0)` -- OpenCV's own float -> 8-bit
  conversion (round to nearest, saturate), exactly what the reference writes."""
  import cv2
  import numpy as np
  cv2.imwrite(path, np.ascontiguousarray(img[:, :, ::-1]) * 255.0)


def steps_montage(trajectory, decisions, operations, masks):
  """The `<name>.steps.png` figure of net.py:843-877: row 0 the 64x64 trajectory (input + one image per
  step), below it -- shifted half a cell, between consecutive images -- each step's decision (pdf), operation
  (the selected filter's drawing) and mask panels from the agent's debugger."""
  import cv2
  import numpy as np
  padding, patch = 4, 64
  grid = patch + padding
  steps = len(trajectory)
  fused = np.ones(shape=(grid * 4, grid * steps, 3), dtype=np.float32)
  for i, im in enumerate(trajectory):
    fused[0:patch, grid * i:grid * i + patch] = cv2.resize(im, dsize=(patch, patch), interpolation=cv2.INTER_NEAREST)
  for i in range(steps - 1):
    sx = grid * i + grid // 2
    for sy, panel in ((grid, decisions[i]), (grid * 2 - padding // 2, operations[i]), (grid * 3 - padding, masks[i])):
      fused[sy:sy + patch, sx:sx + patch] = cv2.resize(panel, dsize=(patch, patch), interpolation=cv2.INTER_NEAREST)
  return fused


def evaluate_files(trainer, files, output_dir="./outputs", generator=None, step_by_step=False, show_linear=True,
                   show_input=True, debug=True):
  """GAN.eval (net.py:711-877) on a list of image files; images of equal resolution are retouched as one
  batch (evaluate.py:14-18 lists batching as a TODO).  Per input `<name>` it writes what the reference
  writes: `<name>.retouched.png`; `<name>.linear.png` (show_linear); `<name>.input_tone_mapped.png`
  (show_input: max to white, gamma 1/2.4); `<name>.intermediateNN.png` for every step but the last
  (step_by_step); and with debug=True `<name>.steps.png` (steps_montage) and `<name>_debug.pkl` (the list of
  per-step debug_info dicts).  debug=False and step_by_step=False keep the single fused full-resolution
  kernel; otherwise the steps are applied one launch at a time.  Returns {file: ids per step}."""
  import os
  import pickle
  import numpy as np
  from .visualize import make_debugger
  os.makedirs(output_dir, exist_ok=True)
  groups = {}
  for fn in files:
    im = load_linear_image(fn)
    groups.setdefault(im.shape[:2], []).append((fn, im))
  result = {}
  trace = bool(step_by_step or debug)
  for res, items in groups.items():
    batch = torch.from_numpy(np.stack([im for _, im in items])).to(trainer.device)
    out = retouch(trainer, batch, generator=generator, trace=trace)
    o = out["output"].cpu().numpy()
    S = out["ids"].shape[0]
    if trace:
      traj = [t.cpu().numpy() for t in out["trajectory"]]
      inter = [t.cpu().numpy() for t in out["intermediate"]]
      cfg = trainer.cfg
      debugger = make_debugger([cls(out["trajectory"][0], cfg) for cls in cfg.filters], cfg.source_img_size)
    for k, (fn, linear) in enumerate(items):
      base = os.path.join(output_dir, os.path.basename(fn))
      if step_by_step:
        for s in range(S - 1):                                      # the loop breaks before saving the last step
          save_png(base + ".intermediate%02d.png" % s, inter[s][k])
      if show_linear:
        save_png(base + ".linear.png", linear)
      if show_input:
        save_png(base + ".input_tone_mapped.png", (linear / linear.max()) ** (1 / 2.4))
      save_png(base + ".retouched.png", o[k])
      if debug:
        infos = [out["debug"][s][k] for s in range(S)]
        for d in infos:
          d["state"] = d["state"].detach().cpu().numpy()
        with open(base + "_debug.pkl", "wb") as fh:
          pickle.dump(infos, fh)
        panels = [debugger(d, combined=False) for d in infos]
        save_png(base + ".steps.png", steps_montage([t[k] for t in traj], [p[0] for p in panels], [p[1] for p in panels],
                                                    [p[2] for p in panels]))
      result[fn] = out["ids"][:, k].tolist()
  return result
