"""CPU oracle for the policy / critic / value networks -- TEST INFRASTRUCTURE ONLY
(see oracle/__init__.py: pinned to the reference's agent.py / critics.py executed over a TF-1
API stand-in, tests/test_reference_golden.py; not to TF binaries).

torch-CPU restatement of agent.py:11-37 (feature_extractor), critics.py:6-98 (cnn, critic),
filters.py:28-44 (extract_parameters), pdf_sample_layer.py:5-10, agent.py:41-260 and the
losses of net.py:92-199, with every random draw (dropout masks, z, alpha) passed in
explicitly.  Gradients come from torch autograd (create_graph=True for the WGAN-GP term).
Tensors are NHWC like the reference; weights HWIO / [in,out] like its checkpoint."""
import math

import torch
import torch.nn.functional as Fn


def lrelu(x, leak=0.2):
  """util.py:225-229."""
  f1 = 0.5 * (1 + leak)
  f2 = 0.5 * (1 - leak)
  return f1 * x + f2 * torch.abs(x)


def conv4x4s2(x_nhwc, W_hwio, bias=None):
  """ly.conv2d(kernel_size=4, stride=2, padding=SAME): for even sizes SAME == pad 1 each side."""
  y = Fn.conv2d(x_nhwc.permute(0, 3, 1, 2), W_hwio.permute(3, 2, 0, 1), bias=bias, stride=2, padding=1)
  return y.permute(0, 2, 3, 1)


def enrich(img, vec):
  """util.py:31-36: tile a per-image vector over the pixels and concatenate as channels."""
  B, H, W, _ = img.shape
  return torch.cat([img, vec[:, None, None, :].expand(B, H, W, vec.shape[1])], dim=3)


def cnn(net, weights, biases):
  """agent.py:11-33 / critics.py:6-37: (net - 0.5) -> 4 x [conv4x4s2 + lrelu] -> NHWC flatten."""
  net = net - 0.5
  for W, b in zip(weights, biases):
    net = lrelu(conv4x4s2(net, W, b))
  return net.reshape(net.shape[0], -1)


def fc(x, W, b, act=True):
  y = x @ W + b
  return lrelu(y) if act else y


def critic_stats(images):
  """critics.py:48-76: per-image luminance mean / variance and mean saturation."""
  lum = images[..., 0] * 0.27 + images[..., 1] * 0.67 + images[..., 2] * 0.06 + 1e-5
  luminance = lum.mean(dim=(1, 2))
  contrast = ((lum - luminance[:, None, None]) ** 2).mean(dim=(1, 2))     # tf.nn.moments: population variance
  c = images.clamp(0.0, 1.0)
  i_max = c.amax(dim=3)          # amax/amin split the gradient evenly among ties, like TF's reduce_max
  i_min = c.amin(dim=3)
  sat = (i_max - i_min) / (torch.minimum(i_max + i_min, 2.0 - i_max - i_min) + 1e-2)
  saturation = sat.mean(dim=(1, 2))
  return torch.stack([luminance, contrast, saturation], dim=1)


def critic(images, params, states=None):
  """critics.py:42-98.  params: dict conv_w[4], conv_b[4], fc1_w, fc1_b, fc2_w, fc2_b."""
  stat = critic_stats(images)
  vec = stat if states is None else torch.cat([states, stat], dim=1)
  feat = cnn(enrich(images, vec), params["conv_w"], params["conv_b"])
  h = fc(feat, params["fc1_w"], params["fc1_b"])
  return fc(h, params["fc2_w"], params["fc2_b"], act=False)


def feature_extractor(img, states, conv_w, conv_b, drop_mask):
  """agent.py:11-37 on enrich_image_input(img, states); tf.nn.dropout(keep 0.5) == x * mask / 0.5."""
  feat = cnn(enrich(img, states), conv_w, conv_b)
  return feat * drop_mask / 0.5


def pdf_sample(pdf, u):
  """pdf_sample_layer.py:5-10."""
  pdf = pdf / (pdf.sum(dim=1, keepdim=True) + 1e-36)
  cdf = torch.cumsum(pdf, dim=1) - pdf                      # exclusive cumsum
  return (cdf < u).to(torch.int32).sum(dim=1) - 1


def gradient_penalty(interpolated, params, lam=10.0):
  """net.py:174-187: one-sided penalty on ||d critic / d interpolated||."""
  x = interpolated.detach().clone().requires_grad_(True)
  logit = critic(x, params)
  (g,) = torch.autograd.grad(logit.sum(), [x], create_graph=True)
  norm = torch.sqrt(1e-6 + (g ** 2).sum(dim=(1, 2, 3)))
  return lam * torch.mean(torch.clamp(norm - 1.0, min=0.0) ** 2), norm, g


def policy_head(logits, u, states, is_train, progress, cfg):
  """agent.py:100-122, 208-252: pdf, sampled id, surrogate, entropy, head penalties, new states.
  cfg needs: exploration, test_steps, exploration_penalty, filter_usage_penalty."""
  n = logits.shape[1]
  pdf = torch.softmax(logits, dim=1) + 1e-37
  pdf = pdf * (1 - cfg.exploration) + cfg.exploration * 1.0 / n
  pdf = pdf / (pdf.sum(dim=1, keepdim=True) + 1e-30)
  entropy = (-pdf * torch.log(pdf)).sum(dim=1, keepdim=True)
  random_id = pdf_sample(pdf, u)
  max_id = torch.argmax(pdf, dim=1).to(torch.int32)
  ids = is_train * random_id + (1 - is_train) * max_id
  onehot = torch.zeros_like(pdf)
  valid = ids >= 0
  onehot[valid, ids[valid].long()] = 1.0
  surrogate = (onehot * torch.log(pdf + 1e-10)).sum(dim=1, keepdim=True)
  is_last = (torch.abs(states[:, 2:3] + 1 - cfg.test_steps) < 1e-4).to(logits.dtype)
  usage = states[:, 3:]
  usage_penalty = (usage * onehot).sum(dim=1, keepdim=True)
  new_states = torch.cat([is_last, is_last, states[:, 2:3] + 1, torch.maximum(usage, onehot)], dim=1)
  entropy_penalty = (1.0 - progress) * cfg.exploration_penalty * (-entropy + math.log(n))
  penalty_head = entropy_penalty + usage_penalty * cfg.filter_usage_penalty
  return pdf, ids, surrogate, entropy, penalty_head, new_states


def rl_losses(fake_logit, fake_input_logit, old_value, new_value, penalty, surrogate, new_states, cfg):
  """net.py:92-163 (WGAN branch, use_TD).  All inputs [B,1]; returns (g_loss, v_loss)."""
  stopped = new_states[:, 1:2]
  clear_final = (new_states[:, 2:3] > cfg.maximum_trajectory_length).to(fake_logit.dtype)
  new_value = new_value * (1.0 - clear_final)
  raw_reward = (cfg.all_reward + (1 - cfg.all_reward) * stopped) * \
      (fake_logit - fake_input_logit.detach()) * cfg.critic_logit_multiplier
  reward = raw_reward - penalty if cfg.use_penalty else raw_reward
  q = reward + (1.0 - stopped) * cfg.discount_factor * new_value
  advantage = q.detach() - old_value
  v_loss = (advantage ** 2).mean()
  g_loss = (-q * cfg.parameter_lr_mul + surrogate * (-advantage).detach()).mean()
  return g_loss, v_loss
