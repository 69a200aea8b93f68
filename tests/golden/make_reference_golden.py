#!/usr/bin/env python
"""Golden vectors produced by EXECUTING THE REFERENCE'S OWN PYTHON in the build container.

  python tests/golden/make_reference_golden.py         ->  tests/golden/reference_golden.npz

What runs: /root/reference/{filters,agent,critics,pdf_sample_layer,util,config_example,replay_memory,net}.py, imported
unmodified from where they lie (read-only), over tests/golden/tf1_shim -- an eager stand-in for the
slice of the TensorFlow 1.x API those files call (TF 1.6 is not installable here).  So the op order,
constants, broadcasting, variable scopes and every formula are the reference's; the TF primitives
underneath (exp, pow, clip tie rules, RGB<->HSV functors, moments, SAME conv, dropout) are the shim's
restatement of TF 1.6 on torch CPU, and torch autograd stands in for tf.gradients.  The only source
patch: util.py:658-664 names a variable `async` (a keyword since Python 3.7) inside its self-test,
which is renamed in memory before compiling; `skimage` / `tifffile` (util.py:484-485, image file I/O
only) are stubbed.  net.py's graph construction (GAN.__init__) is executed too (section 5b): its
placeholders are handed the batch a sess.run would feed, so the graph-building code computes eagerly;
section 5 keeps an independent line-by-line restatement of net.py:92-194 next to it.

Sections (all inputs and random draws are stored next to the outputs):
  1. every Filter subclass: filter_param_regressor + process, fp64 and fp32 runs, autograd gradients
  2. Filter.apply with masking on (extract_parameters -> regressor -> get_mask -> lerp) + high_res
  3. agent_generator: 5-step rollouts (argmax and sampled) + the high_res branch
  4. critic / value network
  5. critic loss with the WGAN-GP term and its double-backward; generator / value losses (restated lines)
  5b. the same from net.py's own GAN.__init__ executed eagerly (losses, rewards, tf.gradients per optimizer)
  6. the cv2 visual debugger: every Filter.visualize_filter / visualize_mask, agent_generator's debugger
  7. ReplayMemory: the records every generator / critic batch draws over 40 iterations (replay_memory.py)
Sections 3-5 run twice: with the shipped pretrained checkpoint ("pre_": compact outputs; the weights
do not travel, so only tests in the build container can use them) and with the name-seeded weights of
tests/golden/seeded_weights.py ("seed_": full outputs, reproducible on the GPU box).

/root/reference does not exist on the GPU box: only the .npz travels."""
import os
import re
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, os.path.join(HERE, "tf1_shim"))
sys.path.insert(1, ROOT)
sys.path.insert(2, HERE)

import tensorflow as tf  # noqa: E402  (the shim)
import seeded_weights  # noqa: E402  (tests/golden/seeded_weights.py)

for name in ("skimage", "tifffile"):
  m = types.ModuleType(name)
  m.color = None
  sys.modules[name] = m
src = open(os.path.join(REF, "util.py")).read()
src = re.sub(r"\basync\b", "async_", src)
util = types.ModuleType("util")
util.__file__ = os.path.join(REF, "util.py")
sys.modules["util"] = util
exec(compile(src, util.__file__, "exec"), util.__dict__)
sys.path.insert(3, REF)

import builtins  # noqa: E402

_print = builtins.print
builtins.print = lambda *a, **k: None            # the reference prints shapes while building
import filters as RF  # noqa: E402  (reference)
import agent as RA  # noqa: E402
import critics as RC  # noqa: E402
import config_example as RCFG  # noqa: E402
builtins.print = _print

cfg = RCFG.cfg
OUT = {}
FILTERS = [RF.ExposureFilter, RF.GammaFilter, RF.ImprovedWhiteBalanceFilter, RF.SaturationPlusFilter, RF.ToneFilter,
           RF.ContrastFilter, RF.WNBFilter, RF.ColorFilter, RF.LevelFilter, RF.VignetFilter]
assert cfg.filters == FILTERS[:8]


def quiet(fn, *a, **k):
  builtins.print = lambda *aa, **kk: None
  try:
    return fn(*a, **k)
  finally:
    builtins.print = _print


def images(B, H, W, seed):
  """dark linear-RGB-like pixels with the clamp / knot / tie cases planted."""
  g = torch.Generator().manual_seed(seed)
  x = torch.exp(torch.randn(B, H, W, 3, generator=g, dtype=torch.float64) - 1.6).clamp(0, 4)
  flat = x.reshape(-1)
  n = flat.numel()
  idx = torch.randperm(n, generator=g)
  k = max(n // 50, 8)
  flat[idx[:k]] = 1 + 3 * torch.rand(k, generator=g, dtype=torch.float64)
  flat[idx[k:k + 4]] = 0.0
  flat[idx[k + 4:k + 8]] = 0.001
  for i in range(1, 9):
    flat[idx[k + 8 + 2 * i:k + 10 + 2 * i]] = i / 8.0
  px = x.reshape(-1, 3)
  px[1] = px[1, 0]                                      # grey pixel (hue undefined)
  px[2, 1] = px[2, 0]                                   # two-channel ties (arg-max / arg-min rules)
  px[3, 2] = px[3, 1]
  px[4] = 0.0                                           # black pixel
  return x


def put(prefix, **kw):
  for k, v in kw.items():
    if isinstance(v, torch.Tensor):
      v = v.detach().numpy()
    OUT["%s_%s" % (prefix, k)] = np.asarray(v)


# ---------------------------------------------------------------------------------------------
# 1. filter_param_regressor + process (filters.py:170-498), gradients by autograd
def section_filters():
  B, H, W = 3, 12, 10
  for fid, cls in enumerate(FILTERS):
    x = images(B, H, W, seed=1000 + fid)
    res = {}
    for dt in (torch.float64, torch.float32):
      tf.set_float_dtype(dt)
      img = x.to(dt).clone().requires_grad_(True)
      f = quiet(cls, img, cfg)
      n = f.get_num_filter_parameters()
      g = torch.Generator().manual_seed(2000 + fid)
      logits = torch.randn(B, n, generator=g, dtype=torch.float64).to(dt).requires_grad_(True)
      gy = torch.randn(B, H, W, 3, generator=g, dtype=torch.float64).to(dt)
      param = f.filter_param_regressor(logits)
      param.retain_grad()
      y = f.process(img, param)
      (y * gy).sum().backward()
      zg = lambda g, like: torch.zeros_like(like) if g is None else g          # Vignet.process ignores its parameter
      res[dt] = dict(y=y, param=param, gx=img.grad, gparam=zg(param.grad, param), glogits=zg(logits.grad, logits),
                     logits=logits, gy=gy)
    r64, r32 = res[torch.float64], res[torch.float32]
    put("f%d" % fid, name=f.get_short_name(), x=x, logits=r64["logits"], gy=r64["gy"], param=r64["param"], y=r64["y"],
        gx=r64["gx"], gparam=r64["gparam"], glogits=r64["glogits"], y32=r32["y"], param32=r32["param"],
        gx32=r32["gx"], glogits32=r32["glogits"])
  tf.set_float_dtype(torch.float32)


# ---------------------------------------------------------------------------------------------
# 2. Filter.apply with masking (filters.py:28-148), incl. the high_res branch (filters.py:89-96)
def section_masked():
  B, H, W, HH, HW = 2, 12, 16, 20, 14
  mcfg = util.Dict(dict(cfg))
  mcfg.masking = True
  mcfg.fc1_size = 32
  tf.set_float_dtype(torch.float64)
  for fid, cls in enumerate(FILTERS):
    x = images(B, H, W, seed=3000 + fid).requires_grad_(True)
    hr = images(B, HH, HW, seed=3100 + fid)
    g = torch.Generator().manual_seed(3200 + fid)
    feat = torch.randn(B, 16, generator=g, dtype=torch.float64)
    gy = torch.randn(B, H, W, 3, generator=g, dtype=torch.float64)
    tf.reset_variables()
    tf.initializer[0] = lambda name, shape: seeded_weights.make(name, shape, seed=31)
    f = quiet(cls, x, mcfg)
    with tf.variable_scope("filter_%d" % fid):
      low, high, dbg = quiet(f.apply, x, img_features=feat, high_res=hr)
    V = tf.variables()
    names = tf.created_variables()
    grads = torch.autograd.grad((low * gy).sum(), [x] + [V[k] for k in names])
    put("m%d" % fid, x=x, hr=hr, feat=feat, gy=gy, low=low, high=high, mask=f.mask, high_mask=f.high_res_mask,
        mask_parameters=f.mask_parameters, gx=grads[0])
    for k, gk in zip(names, grads[1:]):
      short = k.split("/", 1)[1].replace("/", "_")
      put("m%d" % fid, **{short: V[k], "g_" + short: gk})
    OUT["m%d_varnames" % fid] = np.array(names)
  tf.set_float_dtype(torch.float32)


# ---------------------------------------------------------------------------------------------
def load_checkpoint():
  from exposure_b200 import tf_bundle
  named = tf_bundle.load_bundle(os.path.join(REF, "models/example/pretrained/model.ckpt-20000"))
  return {k: np.asarray(v) for k, v in named.items() if np.asarray(v).dtype == np.float32 and np.asarray(v).ndim >= 1}


def use_weights(kind, P, prefixes):
  """kind 'pre': variables pre-loaded from the checkpoint (a missing name is an error);
  kind 'seed': created on demand by tests/golden/seeded_weights.make(name, shape, seed=7)."""
  if kind == "pre":
    tf.reset_variables({k: v for k, v in P.items() if k.split("/")[0] in prefixes})

    def missing(name, shape):
      raise KeyError("the reference asked for variable %s which the shipped checkpoint does not hold" % name)
    tf.initializer[0] = missing
  else:
    tf.reset_variables()
    tf.initializer[0] = lambda name, shape: seeded_weights.make(name, shape, seed=7)


def thumbnails():
  """64x64 linear thumbnails of shipped sample inputs the way net.py:726,779 makes them
  (linearize_ProPhotoRGB / get_image_center: util.py:495-501, 311-323, 86-94) + 2 synthetic."""
  import cv2
  out = []
  for n in ("A.tif", "H.tif"):
    im = cv2.imread(os.path.join(REF, "models/sample_inputs", n), cv2.IMREAD_UNCHANGED)[:, :, ::-1] / 65535.0
    lin = util.linearize_ProPhotoRGB(im)
    out.append(cv2.resize(util.get_image_center(lin), dsize=(64, 64)))
  syn = images(2, 64, 64, seed=4000).numpy() * 0.5
  return np.concatenate([np.stack(out), syn]).astype(np.float32)


def compact(t):
  """what is kept of an image batch when the full tensor would not be reproducible elsewhere anyway:
  every 37th value of each image (fp64) -- tests apply the same reduction to their own result."""
  return t.detach().reshape(t.shape[0], -1)[:, ::37].double().clone()


class Draws:
  """Records every tf.nn.dropout draw in call order so the same masks can be replayed elsewhere."""

  def __init__(self, seed):
    self.g = torch.Generator().manual_seed(seed)
    self.log = []

  def __call__(self, shape, keep):
    u = torch.rand(shape, generator=self.g, dtype=torch.float64)
    self.log.append(torch.floor(keep + u).to(torch.uint8))      # keep-mask; the multiplier is mask / keep
    return u


GRAD_NAMES = ["generator/filter_0/fc2/weights", "generator/filter_3/fc2/weights", "generator/filter_4/fc2/weights",
              "generator/filter_7/fc2/biases", "generator/Conv_3/biases", "generator/Conv/weights",
              "generator/action_selection/selector_fc2/weights", "generator/action_selection/Conv_2/biases"]


# 3. agent_generator (agent.py:41-260)
def section_agent(kind, P):
  img0 = torch.from_numpy(thumbnails())
  B = img0.shape[0]
  full = kind == "seed"
  OUT["thumbs"] = img0.numpy()            # the image batch of sections 3-5
  for dt, tag in ((torch.float64, ""), (torch.float32, "32")):
    tf.set_float_dtype(dt)
    for mode, is_train in (("argmax", 0), ("sample", 1)):
      use_weights(kind, P, ("generator",))
      draws = Draws(5000 + is_train)
      tf.nn.dropout_source = draws
      g = torch.Generator().manual_seed(5100 + is_train)
      img = img0.to(dt)
      states = torch.zeros(B, cfg.num_state_dim, dtype=dt)
      for step in range(cfg.test_steps):
        z = torch.rand(B, cfg.z_dim, generator=g, dtype=torch.float64).to(dt)
        n0 = len(draws.log)
        with tf.variable_scope("generator"):
          (net, new_states, surrogate, penalty), dbg, _ = quiet(RA.agent_generator, [img, z, states], is_train=is_train,
                                                                progress=0.25, cfg=cfg)
        pre = "%s_ag_%s_s%d" % (kind, mode, step)
        if tag == "":
          put(pre, z0=z[:, 0], states=states, drop_f=draws.log[n0], drop_s=draws.log[n0 + 1],
              out=net.float() if full else compact(net), new_states=new_states, surrogate=surrogate, penalty=penalty,
              pdf0=dbg["pdf"], id0=dbg["selected_filter_id"])
          if step == 0:
            if kind == "pre":
              assert sorted(tf.created_variables()) == sorted(k for k in P if k.startswith("generator/")), \
                  "variable names differ from the checkpoint"
            else:
              OUT["seed_generator_varnames"] = np.array(tf.created_variables())
              OUT["seed_generator_varshapes"] = np.array([",".join(map(str, tf.variables()[k].shape)) for k in tf.created_variables()])
            # gradient of a fixed functional of the outputs w.r.t. a few variables (tf.gradients stand-in)
            gw = torch.sin(0.37 * torch.arange(net.numel(), dtype=torch.float64)).reshape(net.shape)   # d(loss)/d(out)
            L = (net * gw).sum() + surrogate.sum() * 0.7 + penalty.sum() * 1.3
            gs = torch.autograd.grad(L, [tf.variables()[k] for k in GRAD_NAMES], allow_unused=True)
            for k, gk in zip(GRAD_NAMES, gs):
              put(pre, **{"grad_" + k.replace("/", "."): gk})
        else:
          put(pre, surrogate32=surrogate, penalty32=penalty)
          if step == cfg.test_steps - 1:
            put(pre, out32=net if full else compact(net))
        img, states = net.detach(), new_states.detach()
  # high-resolution branch (agent.py:126-129, 257-260): one step on a non-square full-res batch
  tf.set_float_dtype(torch.float64)
  use_weights(kind, P, ("generator",))
  draws = Draws(5200)
  tf.nn.dropout_source = draws
  hr = images(B, 40, 56, seed=5201) * 0.3
  z = torch.rand(B, cfg.z_dim, generator=torch.Generator().manual_seed(5202), dtype=torch.float64)
  states = torch.zeros(B, cfg.num_state_dim, dtype=torch.float64)
  with tf.variable_scope("generator"):
    (net, new_states, hro), _, _ = quiet(RA.agent_generator, [img0.double(), z, states], is_train=0, progress=0, cfg=cfg,
                                         high_res=hr)
  put(kind + "_ag_hr", hr=hr.float(), z0=z[:, 0], drop_f=draws.log[0], drop_s=draws.log[1],
      out=net.float() if full else compact(net), high=hro.float() if full else compact(hro), new_states=new_states)
  tf.set_float_dtype(torch.float32)


# 4. critic / value (critics.py:42-98; called as in net.py:68-90)
def section_critic(kind, P):
  img = torch.from_numpy(thumbnails())
  g = torch.Generator().manual_seed(6000)
  B = img.shape[0]
  states = torch.zeros(B, cfg.num_state_dim, dtype=torch.float64)
  states[:, 2] = torch.arange(B, dtype=torch.float64) % 5
  states[:, 3:] = (torch.rand(B, 8, generator=g, dtype=torch.float64) < 0.3).double()
  for dt, tag in ((torch.float64, ""), (torch.float32, "32")):
    tf.set_float_dtype(dt)
    use_weights(kind, P, ("critic", "rl_value"))
    x = img.to(dt).clone().requires_grad_(True)
    logit, _, _ = quiet(RC.critic, images=x, cfg=cfg, is_train=False)
    logit2, _, _ = quiet(RC.critic, images=x * 2.0, cfg=cfg, reuse=True, is_train=False)
    with tf.variable_scope("rl_value"):
      value, _, _ = quiet(RC.critic, images=x, states=states.to(dt), cfg=cfg, reuse=False, is_train=False)
    (gimg,) = torch.autograd.grad(logit.sum(), [x])
    if kind == "seed" and tag == "":
      OUT["seed_critic_varnames"] = np.array(tf.created_variables())
      OUT["seed_critic_varshapes"] = np.array([",".join(map(str, tf.variables()[k].shape)) for k in tf.created_variables()])
    put(kind + "_cr", **{"logit" + tag: logit, "logit_x2" + tag: logit2, "value" + tag: value})
    if tag == "":
      put(kind + "_cr", dlogit_dimg=gimg.float() if kind == "seed" else compact(gimg))
  put(kind + "_cr", states=states)
  tf.set_float_dtype(torch.float32)


# 5. losses of net.py:92-194 on top of the reference callables (restated lines cited inline)
def section_losses(kind, P):
  tf.set_float_dtype(torch.float64)
  use_weights(kind, P, ("generator", "critic", "rl_value"))
  draws = Draws(7000)
  tf.nn.dropout_source = draws
  g = torch.Generator().manual_seed(7001)
  fake_input = torch.from_numpy(thumbnails()).double()
  B = fake_input.shape[0]
  real = (images(B, 64, 64, seed=7002) * 0.6).clamp(0, 1)
  z = torch.rand(B, cfg.z_dim, generator=g, dtype=torch.float64)
  states = torch.zeros(B, cfg.num_state_dim, dtype=torch.float64)
  states[:, 2] = torch.tensor([0, 2, 4, 4], dtype=torch.float64)[:B]
  alpha = torch.rand(B, 1, 1, 1, generator=g, dtype=torch.float64)
  progress = 0.4
  with tf.variable_scope("generator"):                                                     # net.py:56-61
    (fake_output, new_states, surrogate, penalty), _, _ = quiet(RA.agent_generator, [fake_input, z, states], is_train=1,
                                                                progress=progress, cfg=cfg)
  if kind == "seed":
    # untrained critic: scale its head so that ||d logit / d image|| > 1 and the one-sided penalty is active
    with torch.no_grad():
      quiet(RC.critic, images=real, cfg=cfg, is_train=True)
      tf.variables()["critic/fully_connected_1/weights"].mul_(40.0)
  real_logit, _, _ = quiet(RC.critic, images=real, cfg=cfg, reuse=kind == "seed", is_train=True)   # net.py:68-73
  fake_logit, _, _ = quiet(RC.critic, images=fake_output, cfg=cfg, reuse=True, is_train=True)
  fake_input_logit, _, _ = quiet(RC.critic, images=fake_input, cfg=cfg, reuse=True, is_train=True)
  with tf.variable_scope("rl_value"):                                                      # net.py:77-90
    old_value, _, _ = quiet(RC.critic, images=fake_input, states=states, cfg=cfg, reuse=False, is_train=True)
    new_value, _, _ = quiet(RC.critic, images=fake_output, states=new_states, cfg=cfg, reuse=True, is_train=True)
  stopped = new_states[:, util.STATE_STOPPED_DIM:util.STATE_STOPPED_DIM + 1]             # net.py:92-97
  clear_final = tf.cast(new_states[:, util.STATE_STEP_DIM:util.STATE_STEP_DIM + 1] > cfg.maximum_trajectory_length, tf.float32)
  new_value = new_value * (1.0 - clear_final)
  raw_reward = (cfg.all_reward + (1 - cfg.all_reward) * stopped) * (
      fake_logit - tf.stop_gradient(fake_input_logit)) * cfg.critic_logit_multiplier        # net.py:108-110
  reward = raw_reward - penalty                                                            # net.py:112-113
  q_value = reward + (1.0 - stopped) * cfg.discount_factor * new_value                    # net.py:125-126
  advantage = tf.stop_gradient(q_value) - old_value                                        # net.py:128
  v_loss = tf.reduce_mean(advantage ** 2, axis=(0, 1))                                     # net.py:129
  c_loss = tf.reduce_mean(fake_logit - real_logit)                                         # net.py:151
  routine_loss = -q_value * cfg.parameter_lr_mul                                           # net.py:153-154
  g_loss = tf.reduce_mean(routine_loss + surrogate * tf.stop_gradient(-advantage))         # net.py:161-162
  interpolated = (real + alpha * (fake_output.detach() - real)).requires_grad_(True)      # net.py:176
  inte_logit, _, _ = quiet(RC.critic, images=interpolated, cfg=cfg, reuse=True, is_train=True)
  gradients = tf.gradients(inte_logit, [interpolated])[0]                                  # net.py:181-183
  gradient_norm = tf.sqrt(1e-6 + tf.reduce_sum(gradients ** 2, axis=[1, 2, 3]))            # net.py:185
  gradient_penalty = cfg.gradient_penalty_lambda * tf.reduce_mean(tf.maximum(gradient_norm - 1.0, 0.0) ** 2)
  c_total = c_loss + gradient_penalty                                                      # net.py:194
  V = tf.variables()
  cnames = sorted(k for k in V if k.startswith("critic/"))
  vnames = sorted(k for k in V if k.startswith("rl_value/"))
  # the critic step differentiates c_loss w.r.t. theta_c with fake_output a fed constant (net.py:362-368)
  fake_c, _, _ = quiet(RC.critic, images=fake_output.detach(), cfg=cfg, reuse=True, is_train=True)
  c_step = tf.reduce_mean(fake_c - real_logit) + gradient_penalty
  gc = torch.autograd.grad(c_step, [V[k] for k in cnames], retain_graph=True)
  gv = torch.autograd.grad(v_loss, [V[k] for k in vnames], retain_graph=True, allow_unused=True)
  gg = torch.autograd.grad(g_loss, [V[k] for k in GRAD_NAMES], retain_graph=True, allow_unused=True)
  pre = kind + "_ls"
  put(pre, real=real.float(), z0=z[:, 0], states=states, alpha=alpha.reshape(-1), progress=progress,
      drop_f=draws.log[0], drop_s=draws.log[1], fake_output=fake_output.float() if kind == "seed" else compact(fake_output),
      new_states=new_states, surrogate=surrogate, penalty=penalty, g_loss=g_loss, v_loss=v_loss, c_loss=c_total, emd=-c_loss,
      gradient_penalty=gradient_penalty, critic_gradient_norm=tf.reduce_mean(gradient_norm), fake_logit=fake_logit,
      real_logit=real_logit, old_value=old_value, new_value=new_value)
  for k, gk in list(zip(cnames, gc)) + list(zip(vnames, gv)) + list(zip(GRAD_NAMES, gg)):
    small = k.endswith("biases") or "fully_connected_1" in k or "fc2" in k or k.endswith("Conv/weights")
    if gk is None:
      gk = torch.zeros_like(V[k])
    if small:
      put(pre, **{"grad_" + k.replace("/", "."): gk})
    else:                                    # big tensors: keep a strided sample + the exact norm
      put(pre, **{"gradsample_" + k.replace("/", "."): gk.reshape(-1)[::997].clone(), "gradnorm_" + k.replace("/", "."): gk.norm()})
  tf.set_float_dtype(torch.float32)


# 5b. the same losses from net.py ITSELF: GAN.__init__ (net.py:20-284) builds its whole graph -- generator,
#     critic x3, value x2, rewards, TD target, losses, gradient penalty, the three ly.optimize_loss -- and runs
#     eagerly here because every tf.placeholder already holds the batch a sess.run would feed (tf.feeds).
#     ly.optimize_loss records tf.gradients(loss, theta); nothing is stepped.
def section_net_graph(kind, P):
  import tempfile
  import tensorflow.contrib.layers as ly
  import net as RN  # noqa: E402  (reference net.py)
  tf.set_float_dtype(torch.float64)
  use_weights(kind, P, ("generator", "critic", "rl_value"))
  draws = Draws(7000)
  tf.nn.dropout_source = draws
  g = torch.Generator().manual_seed(7001)
  fake_input = torch.from_numpy(thumbnails()).double()
  B = fake_input.shape[0]
  real = (images(B, 64, 64, seed=7002) * 0.6).clamp(0, 1)
  z = torch.rand(B, cfg.z_dim, generator=g, dtype=torch.float64)
  states = torch.zeros(B, cfg.num_state_dim, dtype=torch.float64)
  states[:, 2] = torch.tensor([0, 2, 4, 4], dtype=torch.float64)[:B]
  alpha = torch.rand(B, 1, 1, 1, generator=g, dtype=torch.float64)          # same draws as section 5
  progress = 0.4
  if kind == "seed":
    with torch.no_grad():
      quiet(RC.critic, images=real, cfg=cfg, is_train=True)
      tf.variables()["critic/fully_connected_1/weights"].mul_(40.0)
    tf._store.counts.clear()                                                # as if the graph were built from scratch
  tf.contrib.distributions.source = lambda shape: alpha.reshape(shape)
  tf.contrib.distributions.log.clear()
  ly.optimize_log.clear()
  tf.feeds.clear()
  tf.feeds.update(fake_input=fake_input, real_data=real, z=z, states=states, progress=torch.tensor(progress, dtype=torch.float64),
                  is_train=torch.tensor(1, dtype=torch.int32), lr_g=torch.tensor(1e-5, dtype=torch.float64),
                  lr_c=torch.tensor(1e-5, dtype=torch.float64))
  c2 = util.Dict(dict(cfg))
  c2.name = "golden"
  c2.batch_size = B                                                         # alpha's shape (net.py:176)
  c2.real_data_provider = lambda: None
  cwd = os.getcwd()
  with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    try:
      gan = quiet(RN.GAN, c2, restore=True)
    finally:
      os.chdir(cwd)
      tf.feeds.clear()
  assert len(ly.optimize_log) == 3 and len(tf.contrib.distributions.log) == 1
  opt_v, opt_g, opt_c = ly.optimize_log                                      # net.py:218, 231, 244 in this order
  pre = kind + "_ng"
  gp = gan.c_loss - tf.reduce_mean(gan.fake_logit - gan.real_logit)          # c_loss = emd part + penalty (net.py:151, 194)
  put(pre, g_loss=gan.g_loss, v_loss=gan.v_loss, c_loss=gan.c_loss, emd=gan.emd, gradient_penalty=gp,
      critic_gradient_norm=gan.critic_gradient_norm, fake_logit=gan.fake_logit, real_logit=gan.real_logit,
      old_value=gan.old_value, new_value=gan.new_value, new_states=gan.new_states, surrogate=gan.surrogate_loss_addition,
      penalty=gan.penalty, reward=gan.reward, q_value=gan.q_value, advantage=gan.advantage,
      fake_output=gan.fake_output.float() if kind == "seed" else compact(gan.fake_output),
      n_theta=np.array([len(gan.theta_g), len(gan.theta_v), len(gan.theta_c)]))
  for rec in (opt_v, opt_g, opt_c):
    for k, gk in rec["grads"].items():
      small = k.endswith("biases") or "fully_connected_1" in k or "fc2" in k or k.endswith("Conv/weights")
      if k.startswith("generator/") and k not in GRAD_NAMES:
        continue
      if small:
        put(pre, **{"grad_" + k.replace("/", "."): gk})
      else:
        put(pre, **{"gradsample_" + k.replace("/", "."): gk.reshape(-1)[::997].clone(), "gradnorm_" + k.replace("/", "."): gk.norm()})
  tf.set_float_dtype(torch.float32)


# 6. visual debugger: Filter.visualize_filter / visualize_mask (filters.py:150-168 + overrides) and the
#    debugger closure of agent_generator (agent.py:141-204) -- host-side cv2 drawing on debug_info
def section_visualize(P):
  import cv2
  if not hasattr(cv2, "cv2"):
    cv2.cv2 = cv2            # filters.py:158 spells the constant cv2.cv2.INTER_NEAREST (old OpenCV wheels)
  tf.set_float_dtype(torch.float32)
  dummy = torch.zeros(1, 64, 64, 3)
  g = np.random.RandomState(8000)
  mask = g.rand(12, 16, 1).astype(np.float32)
  for fid, cls in enumerate(FILTERS):
    f = quiet(cls, dummy, cfg)
    logits = torch.from_numpy(g.randn(1, f.get_num_filter_parameters()).astype(np.float32))
    param = f.filter_param_regressor(logits)
    dbg = {"filter_parameters": (param if f.debug_info_batched() else param[0]).numpy(), "mask": mask}
    for size in (64, 256):
      canvas = np.full((size, size, 3), 0.5, dtype=np.float32)
      import warnings
      with warnings.catch_warnings():
        warnings.simplefilter("ignore", DeprecationWarning)      # '%.2f' % one-element array (filters.py:210)
        f.visualize_filter(dbg, canvas)
      put("vz%d" % fid, **{"canvas%d" % size: canvas})
    put("vz%d" % fid, param=dbg["filter_parameters"], maskimg=f.visualize_mask(dbg, (64, 64)))
  OUT["vz_mask"] = mask
  # the agent's debugger on one pretrained argmax step and one sampled step
  img0 = torch.from_numpy(thumbnails())
  for mode, is_train in (("argmax", 0), ("sample", 1)):
    use_weights("pre", P, ("generator",))
    tf.nn.dropout_source = Draws(8100 + is_train)
    z = torch.rand(img0.shape[0], cfg.z_dim, generator=torch.Generator().manual_seed(8200 + is_train))
    states = torch.zeros(img0.shape[0], cfg.num_state_dim)
    with tf.variable_scope("generator"):
      _, dbg, debugger = quiet(RA.agent_generator, [img0, z, states], is_train=is_train, progress=0.5, cfg=cfg)
    host = {"state": dbg["state"].numpy(), "selected_filter_id": int(dbg["selected_filter_id"]),
            "pdf": dbg["pdf"].detach().numpy(),
            "filter_debug_info": [{"filter_parameters": d["filter_parameters"].detach().numpy(), "mask": d["mask"].detach().numpy()}
                                  for d in dbg["filter_debug_info"]]}
    combined = quiet(debugger, host, combined=True)
    panels = quiet(debugger, host, combined=False)
    pre = "vzdbg_" + mode
    put(pre, selected=host["selected_filter_id"], pdf=host["pdf"], combined=combined, panel_pdf=panels[0],
        panel_detail=panels[1], panel_mask=panels[2], width=debugger.width)
    for j, d in enumerate(host["filter_debug_info"]):
      put(pre, **{"param%d" % j: d["filter_parameters"], "mask%d" % j: d["mask"]})


# 7. ReplayMemory (replay_memory.py:8-282): which records each generator / critic batch draws, driven the way
#    GAN.train drives it (net.py:325-362) with counting stand-ins for the data providers.  All randomness is
#    Python's `random` (shuffle / random()), seeded.
def section_replay():
  import random
  import replay_memory as RM  # noqa: E402  (reference)

  class Counting:
    """data-provider stand-in: image k is filled with the value k (so a record's identity survives)."""

    def __init__(self, start):
      self.n = start

    def get_next_batch(self, bs):
      ids = np.arange(self.n, self.n + bs, dtype=np.float32)
      self.n += bs
      return np.tile(ids[:, None, None, None], (1, 4, 4, 3)), np.zeros(bs, dtype=np.float32)

  for tag, test_steps in (("a", 5), ("b", 9)):               # b: trajectories outlive maximum_trajectory_length = 7
    c2 = util.Dict(dict(cfg))
    c2.source_img_size = c2.real_img_size = 4
    c2.replay_memory_size, c2.batch_size, c2.test_steps = 8, 4, test_steps     # pool = 2 x batch like config_example.py:53,131
    c2.fake_data_provider = lambda: Counting(0)
    c2.fake_data_provider_test = lambda: Counting(100000)
    c2.real_data_provider = lambda: Counting(200000)
    random.seed(4242)
    mem = RM.ReplayMemory(c2, load=True)
    B = c2.batch_size
    kinds, ids, steps = [], [], []
    for it in range(40):
      images, states, features = mem.get_next_fake_batch(B)                       # net.py:326 -> replay_memory.py:230-246
      kinds.append(0); ids.append(images[:, 0, 0, 0].copy()); steps.append(states[:, util.STATE_STEP_DIM].copy())
      new_states = states.copy()                                                   # agent.py:208-222
      last = (np.abs(states[:, util.STATE_STEP_DIM] + 1 - c2.test_steps) < 1e-4).astype(np.float32)
      new_states[:, util.STATE_REWARD_DIM] = last
      new_states[:, util.STATE_STOPPED_DIM] = last
      new_states[:, util.STATE_STEP_DIM] = states[:, util.STATE_STEP_DIM] + 1
      mem.replace_memory(mem.images_and_states_to_records(images, new_states, features))   # net.py:340-342
      # the reference asserts when the pool holds no terminated record (hence its 100 warm-up generator steps,
      # net.py:318-320); the driver here simply skips the critic batches of such an iteration
      if any(r.state[util.STATE_STOPPED_DIM] > 0 for r in mem.image_pool):
        for _ in range(2):
          images, states, features = mem.replay_fake_batch(B)                    # net.py:357 -> replay_memory.py:249-273
          kinds.append(1); ids.append(images[:, 0, 0, 0].copy()); steps.append(states[:, util.STATE_STEP_DIM].copy())
    put("rp_" + tag, kinds=np.array(kinds, dtype=np.int32), ids=np.stack(ids).astype(np.int64), steps=np.stack(steps).astype(np.int32),
        test_steps=test_steps, seed=4242, pool=8, batch=4)


def main():
  torch.set_num_threads(max(1, os.cpu_count() or 1))
  section_filters()
  section_masked()
  P = load_checkpoint()
  for kind in ("pre", "seed"):
    section_agent(kind, P)
    section_critic(kind, P)
    section_losses(kind, P)
    section_net_graph(kind, P)
  section_visualize(P)
  section_replay()
  OUT["provenance"] = np.array(
      "reference Python (yuanming-hu/exposure @ 7bb838a: filters.py, agent.py, critics.py, pdf_sample_layer.py, util.py, "
      "config_example.py) executed over tests/golden/tf1_shim (TF-1 API stand-in on torch CPU); not TensorFlow binaries")
  path = os.path.join(HERE, "reference_golden.npz")
  np.savez_compressed(path, **OUT)
  print(path, os.path.getsize(path), "bytes,", len(OUT), "arrays")


if __name__ == "__main__":
  main()
