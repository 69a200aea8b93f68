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
  ids, logits = [], []
  for _ in range(S):
    noise, drop_f, drop_s, _ = trainer.draw(B, generator)
    c = trainer.policy.forward(thumb, states, noise, drop_f, drop_s, 0, 0.0, cfg)
    ids.append(c.ids)
    logits.append(c.logits_sel)
    thumb, states = c.out, c.new_states
  ids = torch.stack(ids).contiguous()
  logits = torch.stack(logits).contiguous()
  if fused:
    out = ops.filter_chain_fwd(high_res.contiguous(), logits, ids, logits=True)
  else:
    out = high_res.contiguous()
    for s in range(S):
      p = ops.filter_regress_fwd(logits[s], ids[s])
      out = ops.filter_fwd(out, p, ids[s])
  return dict(output=out, ids=ids, logits=logits, thumbnails=thumb, states=states)
