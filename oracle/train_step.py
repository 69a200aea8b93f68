"""CPU oracle of the train step -- TEST INFRASTRUCTURE ONLY (parity: see oracle/__init__.py --
pinned to the reference's Python over a TF-1 API stand-in, tests/test_reference_golden.py).

torch-CPU restatement of the generator+value step and the critic step of net.py:56-199 with
autograd supplying every gradient (tf.gradients), executed the way the reference executes it:
all 8 filters evaluated for every image, stacked and one-hot selected (agent.py:58-77,124-125).
Random draws (z, dropout masks, alpha) are inputs.  Parameters are dicts of torch tensors keyed
by the reference's checkpoint variable names."""
import math

import torch

from . import filters as OF
from . import nets as ON

CONV = ["Conv", "Conv_1", "Conv_2", "Conv_3"]


def _stack(P, scope):
  return [P["%s/%s/weights" % (scope, c)] for c in CONV], [P["%s/%s/biases" % (scope, c)] for c in CONV]


def critic_params(P, scope):
  w, b = _stack(P, scope)
  return dict(conv_w=w, conv_b=b, fc1_w=P[scope + "/fully_connected/weights"], fc1_b=P[scope + "/fully_connected/biases"],
              fc2_w=P[scope + "/fully_connected_1/weights"], fc2_b=P[scope + "/fully_connected_1/biases"])


def agent_generator(Pg, img, states, noise, drop_f, drop_s, is_train, progress, cfg, high_res=None):
  """agent.py:41-260 -> (out, new_states, surrogate, penalty, ids, pdf).  drop_* are the
  tf.nn.dropout multipliers (0 or 1/keep) on the NHWC-flattened 4096 features.  With `high_res`
  (agent.py:41, 126-129; filters.py:89-96) every filter is also applied to the full-resolution batch with the
  same parameters and one-hot selected: `out` becomes the pair (out, high_res_out).  No masking on that path."""
  B = img.shape[0]
  w, b = _stack(Pg, "generator")
  feat = ON.cnn(ON.enrich(img, states), w, b) * drop_f.reshape(B, -1)
  filtered, filtered_hi = [], []
  for j in range(8):
    n = OF.NUM_PARAMS[j]
    h = ON.fc(feat, Pg["generator/filter_%d/fc1/weights" % j], Pg["generator/filter_%d/fc1/biases" % j])
    o = ON.fc(h, Pg["generator/filter_%d/fc2/weights" % j], Pg["generator/filter_%d/fc2/biases" % j], act=False)
    if getattr(cfg, "masking", False):      # filters.py:62-148: lerp with the 6-parameter spatial mask
      filtered.append(OF.apply_masked(j, img, o[:, :n], o[:, n:], True, maximum_sharpness=cfg.maximum_sharpness,
                                      minimum_strength=cfg.minimum_strength))
    else:
      filtered.append(OF.apply_filter(j, img, o[:, :n]))
      if high_res is not None:
        filtered_hi.append(OF.apply_filter(j, high_res, o[:, :n]))
  filtered = torch.stack(filtered, dim=1)
  w, b = _stack(Pg, "generator/action_selection")
  sfeat = ON.cnn(ON.enrich(img, states), w, b) * drop_s.reshape(B, -1)
  hs = ON.fc(sfeat, Pg["generator/action_selection/selector_fc1/weights"], Pg["generator/action_selection/selector_fc1/biases"])
  logits = ON.fc(hs, Pg["generator/action_selection/selector_fc2/weights"], Pg["generator/action_selection/selector_fc2/biases"], act=False)
  pdf, ids, surrogate, entropy, pen_head, new_states = ON.policy_head(logits, noise[:, None], states, is_train, progress, cfg)
  onehot = torch.zeros(B, 8, dtype=img.dtype)
  valid = ids >= 0
  onehot[valid, ids[valid].long()] = 1.0
  out = (filtered * onehot[:, :, None, None, None]).sum(dim=1)
  penalty = (torch.clamp(out - 1, min=0) ** 2).mean(dim=(1, 2, 3))[:, None] + pen_head
  if high_res is not None:
    out = (out, (torch.stack(filtered_hi, dim=1) * onehot[:, :, None, None, None]).sum(dim=1))
  return out, new_states, surrogate, penalty, ids, pdf


def generator_step(Pg, Pv, Pc, img, states, noise, drop_f, drop_s, progress, cfg, is_train=1):
  """Losses and gradients of net.py:330's sess.run.  Returns dict with g_loss, v_loss, grads_g
  (dict by name), grads_v, fake_output, new_states."""
  Pg = {k: v.detach().clone().requires_grad_(True) for k, v in Pg.items()}
  Pv = {k: v.detach().clone().requires_grad_(True) for k, v in Pv.items()}
  out, new_states, surrogate, penalty, ids, pdf = agent_generator(Pg, img, states, noise, drop_f, drop_s, is_train, progress, cfg)
  cp = critic_params(Pc, "critic")
  vp = critic_params(Pv, "rl_value/critic")
  fake_logit = ON.critic(out, cp)
  fake_input_logit = ON.critic(img, cp)
  old_value = ON.critic(img, vp, states=states)
  new_value = ON.critic(out, vp, states=new_states)
  g_loss, v_loss = ON.rl_losses(fake_logit, fake_input_logit, old_value, new_value, penalty, surrogate, new_states, cfg)
  names_g = sorted(Pg)
  grads_g = torch.autograd.grad(g_loss, [Pg[k] for k in names_g], retain_graph=True, allow_unused=True)
  names_v = sorted(Pv)
  grads_v = torch.autograd.grad(v_loss, [Pv[k] for k in names_v], allow_unused=True)
  z = lambda g, p: torch.zeros_like(p) if g is None else g
  return dict(g_loss=g_loss.detach(), v_loss=v_loss.detach(), fake_output=out.detach(), new_states=new_states.detach(),
              ids=ids, grads_g={k: z(g, Pg[k]) for k, g in zip(names_g, grads_g)},
              grads_v={k: z(g, Pv[k]) for k, g in zip(names_v, grads_v)},
              fake_logit=fake_logit.detach(), old_value=old_value.detach(), new_value=new_value.detach())


def critic_step(Pc, real, fake, alpha, cfg):
  """net.py:68-71,151,174-194: c_loss = mean(D(fake) - D(real)) + lambda mean(max(||grad||-1,0)^2)."""
  Pc = {k: v.detach().clone().requires_grad_(True) for k, v in Pc.items()}
  cp = critic_params(Pc, "critic")
  real_logit = ON.critic(real, cp)
  fake_logit = ON.critic(fake, cp)
  interpolated = real + alpha[:, None, None, None] * (fake - real)
  gp, norm, _ = ON.gradient_penalty(interpolated, cp, lam=cfg.gradient_penalty_lambda)
  c_loss = (fake_logit - real_logit).mean() + gp
  names = sorted(Pc)
  grads = torch.autograd.grad(c_loss, [Pc[k] for k in names])
  return dict(c_loss=c_loss.detach(), emd=-(fake_logit - real_logit).mean().detach(), gradient_penalty=gp.detach(),
              critic_gradient_norm=norm.mean().detach(), grads_c=dict(zip(names, grads)))


def adam_update(p, g, m, v, lr, t, b1=0.5, b2=0.9, eps=1e-8):
  """tf.train.AdamOptimizer update (config_example.py:158)."""
  lr_t = lr * math.sqrt(1 - b2 ** t) / (1 - b1 ** t)
  m = b1 * m + (1 - b1) * g
  v = b2 * v + (1 - b2) * g * g
  return p - lr_t * m / (v.sqrt() + eps), m, v
