"""Train-step driver: the generator+value step and the WGAN-GP critic step of net.py:298-403 as
explicit kernel schedules, with one all-reduce per optimizer over NVLink (data parallel).

Reference mapping
  generator_step  == sess.run([(opt_g, opt_v), g_loss, v_loss, fake_output, new_states])  net.py:330
  critic_step     == sess.run([opt_c, emd, critic_gradient_norm])                         net.py:362
Losses / rewards follow net.py:92-199 (WGAN branch, use_TD), optimizers config_example.py:134-161
(Adam beta1 0.5, beta2 0.9, lr_g = 0.3*5e-5*0.1^(3t/T), lr_v = 10 lr_g, lr_c = 5e-5*0.1^(3t/T))."""
import math

import torch
import torch.distributed as dist

from . import nn_ops as K
from . import dp as DP
from .nets import CriticNet, ParamArena, ParamStore, PolicyNet


def default_cfg():
  """The hot-path subset of config_example.py (same key names, same values)."""
  from .util import Dict
  from . import filters as _f
  cfg = Dict()
  cfg.filters = [_f.ExposureFilter, _f.GammaFilter, _f.ImprovedWhiteBalanceFilter, _f.SaturationPlusFilter,
                 _f.ToneFilter, _f.ContrastFilter, _f.WNBFilter, _f.ColorFilter]       # config_example.py:22-25
  cfg.curve_steps = 8; cfg.gamma_range = 3; cfg.exposure_range = 3.5
  cfg.color_curve_range = (0.90, 1.10); cfg.tone_curve_range = (0.5, 2)
  cfg.masking = False; cfg.minimum_strength = 0.3; cfg.maximum_sharpness = 1; cfg.clamp = False
  cfg.critic_logit_multiplier = 0.05; cfg.discount_factor = 1.0; cfg.filter_usage_penalty = 1.0
  cfg.use_TD = True; cfg.replay_memory_size = 128; cfg.maximum_trajectory_length = 7
  cfg.over_length_keep_prob = 0.5; cfg.all_reward = 1.0; cfg.img_include_states = True
  cfg.exploration = 0.05; cfg.exploration_penalty = 0.05; cfg.early_stop_penalty = 1.0
  cfg.source_img_size = 64; cfg.base_channels = 32; cfg.dropout_keep_prob = 0.5
  cfg.shared_feature_extractor = True; cfg.fc1_size = 128; cfg.feature_extractor_dims = 4096
  cfg.use_penalty = True; cfg.gan = "w"; cfg.giters = 1; cfg.gradient_penalty_lambda = 10
  cfg.citers = 5; cfg.critic_initialization = 10; cfg.num_state_dim = 11; cfg.z_dim = 3 + 8 * 16
  cfg.test_steps = 5; cfg.real_img_size = 64; cfg.supervised = False; cfg.batch_size = 64
  cfg.max_iter_step = 20000; cfg.parameter_lr_mul = 1; cfg.value_lr_mul = 10
  cfg.lr_g = lambda t: 0.3 * 5e-5 * 0.1 ** (1.0 * t * 3 / cfg.max_iter_step)
  cfg.lr_c = lambda t: 1 * 5e-5 * 0.1 ** (1.0 * t * 3 / cfg.max_iter_step)
  cfg.adam_beta1 = 0.5; cfg.adam_beta2 = 0.9
  return cfg


class Trainer:
  """Owns theta_g / theta_v / theta_c (net.py:205-210), their Adam state and the step schedules."""

  def __init__(self, cfg=None, device=None, seed=0):
    self.cfg = cfg or default_cfg()
    self._check_cfg(self.cfg)
    from . import _cabi
    _cabi.set_filter_ranges(self.cfg)      # cfg.exposure_range / gamma_range / *_curve_range -> the kernels' regressors
    self.device = device or torch.device("cuda", torch.cuda.current_device())
    self.gen = ParamStore(self.device)
    self.policy = PolicyNet(self.gen, n_states=self.cfg.num_state_dim, scope="generator")
    self.val = ParamStore(self.device)
    self.value = CriticNet(self.val, "rl_value/critic", n_states=self.cfg.num_state_dim)
    # theta_g and theta_v are trained by the same sess.run (net.py:330-331): one flat buffer, one gradient exchange
    self.gv = ParamArena(self.device, [self.gen, self.val], [seed, seed + 1])
    self.cri = ParamStore(self.device)
    self.critic = CriticNet(self.cri, "critic", n_states=0)
    self.cri.finalize(seed + 2)
    self.counter_g = self.counter_v = self.counter_c = 0        # net.py:216-241 global steps
    self._g_logit_cache = {}
    # tf.train.ExponentialMovingAverage(decay=0.99, zero_debias=True) of c_average (net.py:119-120, 165-168;
    # updated by every critic step, net.py:268-269): [debiased value, biased accumulator, local_step] on the
    # device -- the checkpoint's mul_8/ExponentialMovingAverage{,/biased,/local_step}
    self.ema_state = torch.zeros(3, device=self.device)
    self._hyper = {k: torch.zeros(1, device=self.device) for k in "gvc"}
    self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    # data-parallel transport (exposure_b200/dp.py): peer-memory all-reduce fused with Adam where the process group
    # allows it (one kernel, graph-capturable), dist.all_reduce + Adam otherwise
    self._peer = None
    if self.world > 1 and DP.peer_exchange_possible(self.device):
      self._peer = {"gv": DP.PeerExchange(self.gv.grad, self.device), "c": DP.PeerExchange(self.cri.grad, self.device)}

  @staticmethod
  def _check_cfg(cfg):
    """Config branches of the reference that this path does not implement fail loudly instead of training
    something else: every shipped config (config_example.py, config_sintel.py) uses the values accepted here."""
    g = lambda k, d: cfg[k] if k in cfg else d
    unsupported = []
    if g("gan", "w") != "w":
      unsupported.append("cfg.gan == %r (LSGAN branch net.py:99-100,131-149; 'not supported' in config_example.py:95)" % cfg["gan"])
    if g("supervised", False):
      unsupported.append("cfg.supervised (net.py:96-97, 212-214)")
    if not g("use_TD", True):
      unsupported.append("cfg.use_TD == False (net.py:156-158: greedy single-step reward)")
    if g("clamp", False):
      unsupported.append("cfg.clamp (agent.py:235-236: clip the filtered image to [0, 5])")
    if not g("shared_feature_extractor", True):
      unsupported.append("cfg.shared_feature_extractor == False (agent.py:62-65: one CNN per filter)")
    if not g("img_include_states", True):
      unsupported.append("cfg.img_include_states == False (util.py:31-36)")
    if g("gradient_penalty_lambda", 10) <= 0:
      unsupported.append("cfg.gradient_penalty_lambda <= 0 (weight clamping, net.py:252-264)")
    # the explicit schedules of nets.py are laid out for the shipped filter list and network sizes: the action
    # index -> kernel id mapping, the fc2 widths and the state vector all follow cfg.filters (agent.py:43-59,
    # util.py:13-16), so anything else is refused instead of silently training the shipped configuration
    from . import filters as _f
    if "filters" in cfg:
      ids = [_f.FILTER_IDS.get(c, None) for c in cfg["filters"]]
      if ids != list(range(8)):
        unsupported.append("cfg.filters must be the shipped list [Exposure, Gamma, ImprovedWhiteBalance, SaturationPlus, Tone, "
                           "Contrast, WNB, Color] in that order (config_example.py:22-25); got kernel ids %s" % ids)
      if g("num_state_dim", 3 + len(ids)) != 3 + len(ids):
        unsupported.append("cfg.num_state_dim must be 3 + len(cfg.filters) (config_example.py:63)")
    for key, want in (("base_channels", 32), ("fc1_size", 128), ("feature_extractor_dims", 4096), ("source_img_size", 64),
                      ("real_img_size", 64), ("curve_steps", 8), ("dropout_keep_prob", None)):
      if want is not None and g(key, want) != want:
        unsupported.append("cfg.%s == %r (the kernels and schedules are built for %r)" % (key, cfg[key], want))
    if unsupported:
      raise NotImplementedError("exposure_b200 implements the shipped configuration of the hot path only; unsupported: "
                                + "; ".join(unsupported))

  @property
  def ema(self):
    v = self.ema_state.tolist()
    return {"value": v[0], "biased": v[1], "local_step": v[2]}

  @ema.setter
  def ema(self, d):
    self.ema_state.copy_(torch.tensor([d["value"], d["biased"], d["local_step"]], dtype=torch.float32))

  def _ema_update(self, c_average, decay=0.99):
    """moving_averages._zero_debias: biased -= (biased - x)(1 - decay); step += 1; value = biased / (1 - decay^step)."""
    st = self.ema_state
    st[1:2].mul_(decay).add_(c_average.reshape(1), alpha=1.0 - decay)
    st[2:3].add_(1.0)
    st[0:1].copy_(st[1:2] / (1.0 - torch.pow(torch.full_like(st[2:3], decay), st[2:3])))

  # ---- optimizer --------------------------------------------------------------------------
  def _set_lr(self, key, lr, t):
    """lr_t of tf.train.AdamOptimizer into the device scalar the fused Adam reads (so that a
    captured graph can be replayed with a new learning rate / step count)."""
    b1, b2 = self.cfg.adam_beta1, self.cfg.adam_beta2
    self._hyper[key].fill_(lr * math.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t))

  def _apply(self, which):
    """The optimizer step of one sess.run: which == "gv" (opt_g and opt_v, net.py:330-331) or "c" (opt_c, net.py:362).
    world > 1: ONE gradient exchange over the flat buffer, the mean (1/world) folded into Adam.
    The lr_t scalars are read from self._hyper (device memory); the whole-iteration graph points the critic's entry at
    a different scalar for each of its critic steps (self._hyper_override)."""
    b1, b2 = self.cfg.adam_beta1, self.cfg.adam_beta2
    hyper = dict(self._hyper, **getattr(self, "_hyper_override", {}))
    if which == "gv":
      buf, parts = self.gv, ((self.gen, "g"), (self.val, "v"))
    else:
      buf, parts = self.cri, ((self.cri, "c"),)
    if self._peer is not None:
      second = hyper[parts[1][1]] if len(parts) > 1 else None
      self._peer[which].allreduce_adam(buf.flat, buf.m, buf.v, hyper[parts[0][1]], parts[0][0].flat.numel(), second, b1, b2)
      return
    if self.world > 1:
      dist.all_reduce(buf.grad)                                  # ONE all-reduce per optimizer step
    for store, key in parts:
      K.adam(store.flat, store.grad, store.m, store.v, hyper[key], b1, b2, 1e-8, 1.0 / self.world)

  # ---- CUDA graphs: each step is a fixed launch sequence, captured once and replayed ---------
  def enable_graphs(self, B=None):
    """Capture the generator+value step and the critic step (forward, backward, all-reduce,
    Adam) into two CUDA graphs over static input buffers.  Per-step scalars (progress, lr_t)
    live in device memory, random draws are made outside and copied in."""
    B = B or self.cfg.batch_size
    dev = self.device
    z = lambda *s: torch.zeros(*s, device=dev)
    self._gi = dict(img=z(B, 64, 64, 3), states=z(B, self.cfg.num_state_dim), noise=z(B), drop_f=z(B, 4, 4, 256),
                    drop_s=z(B, 4, 4, 256), progress=z(1))
    self._ci = dict(real=z(B, 64, 64, 3), fake=z(B, 64, 64, 3), alpha=z(B))
    snap = [t.clone() for s in (self.gv, self.cri) for t in (s.flat, s.m, s.v)]
    ema_snap = self.ema_state.clone()
    for k in "gvc":
      self._hyper[k].zero_()
    # world > 1 with the peer-memory transport: the exchange + Adam is ONE kernel node inside the graph.  With the
    # dist.all_reduce fallback the graphs hold forward + backward only and the all-reduce + Adam are enqueued
    # right after each replay (a process-group collective is not captured).
    self._graph_apply = self.world == 1 or self._peer is not None
    ga = self._graph_apply
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):                                  # warm-up (allocations, lazy attributes)
      for _ in range(2):
        self._generator_impl(self._gi["img"], self._gi["states"], self._gi["noise"], self._gi["drop_f"],
                             self._gi["drop_s"], self._gi["progress"], 1, ga)
        self._critic_impl(self._ci["real"], self._ci["fake"], self._ci["alpha"], ga)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    from . import ops as _ops
    l0 = _ops.launch_count
    self._ggraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self._ggraph):
      self._gout = self._generator_impl(self._gi["img"], self._gi["states"], self._gi["noise"], self._gi["drop_f"],
                                        self._gi["drop_s"], self._gi["progress"], 1, ga)
    l1 = _ops.launch_count
    self._cgraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self._cgraph):
      self._cout = self._critic_impl(self._ci["real"], self._ci["fake"], self._ci["alpha"], ga)
    self.graph_launches = {"generator": l1 - l0, "critic": _ops.launch_count - l1}   # own kernels per replay
    if self.world > 1:
      dist.barrier()                                             # the capture runs ran real exchanges: leave them together
    it = iter(snap)
    for s in (self.gv, self.cri):
      for t in (s.flat, s.m, s.v):
        t.copy_(next(it))
    self.ema_state.copy_(ema_snap)
    self._graph_B = B

  # ---- the WHOLE iteration as one CUDA graph (device replay memory) -----------------------------------------
  def enable_iteration_graph(self, citers=None):
    """Capture one iteration of net.py:307-370 -- draw the generator batch from the device replay memory, random
    draws, generator+value step, re-insert the outputs, then `citers` x (draw a critic batch, critic step) -- into ONE
    CUDA graph.  Needs an attached DeviceReplayMemory and a capturable optimizer step (single GPU, or the peer-memory
    transport).  Per replay the host only stages the provider's fresh RAW / real batches and three lr scalars."""
    from .replay import DeviceReplayMemory
    mem = self.memory
    if not isinstance(mem, DeviceReplayMemory):
      raise TypeError("the whole-iteration graph needs a DeviceReplayMemory (selection logic on the device)")
    if not (self.world == 1 or self._peer is not None):
      raise RuntimeError("the dist.all_reduce transport is not capturable: use the per-step graphs (enable_graphs)")
    cfg, dev = self.cfg, self.device
    B = cfg.batch_size
    n_c = int(citers or cfg.citers)
    s = cfg.source_img_size
    self._it = dict(
        real=torch.zeros(n_c, B, s, s, 3, device=dev),
        hyper=torch.zeros(2 + n_c + 1, device=dev), uni=torch.zeros(B * (1 + n_c), device=dev),
        masks=torch.zeros(2, B, 4, 4, 256, device=dev), citers=n_c)
    I = self._it
    I["progress"] = I["hyper"][2 + n_c:]              # lr_t of the 2 + n_c optimizer steps, then progress: one buffer, one launch
    seed = mem.seed ^ 0x5DEECE66D

    def body():
      img, states = mem.draw_generator()
      K.train_draws(seed, mem.ctl, uniform=I["uni"], mask=I["masks"], keep=float(cfg.dropout_keep_prob))
      self._hyper_override = {"g": I["hyper"][0:1], "v": I["hyper"][1:2]}
      out = self._generator_impl(img, states, I["uni"][:B], I["masks"][0], I["masks"][1], I["progress"], 1, True)
      mem.replace(out["fake_output"], out["new_states"])
      cout = None
      for k in range(n_c):
        fake = mem.draw_critic()
        self._hyper_override = {"c": I["hyper"][2 + k:3 + k]}
        cout = self._critic_impl(I["real"][k], fake, I["uni"][B * (1 + k):B * (2 + k)], True)
      self._hyper_override = {}
      return dict(g_loss=out["g_loss"], v_loss=out["v_loss"], emd=cout["emd"], critic_gradient_norm=cout["critic_gradient_norm"],
                  fake_output=out["fake_output"], new_states=out["new_states"])

    snap = [t.clone() for st in (self.gv, self.cri) for t in (st.flat, st.m, st.v)]
    msnap = [t.clone() for t in (mem.images, mem.states, mem.ctl, self.ema_state)]
    I["hyper"].zero_()
    mem.stage_fresh()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(2):
        body()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    from . import ops as _ops
    l0 = _ops.launch_count
    self._itgraph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self._itgraph):
      self._itout = body()
    self.graph_launches = dict(getattr(self, "graph_launches", {}), iteration=_ops.launch_count - l0)
    if self.world > 1:
      dist.barrier()
    it = iter(snap)
    for st in (self.gv, self.cri):
      for t in (st.flat, st.m, st.v):
        t.copy_(next(it))
    for t, sv in zip((mem.images, mem.states, mem.ctl, self.ema_state), msnap):
      t.copy_(sv)

  def _iteration_replay(self, it):
    cfg, I, mem = self.cfg, self._it, self.memory
    B = cfg.batch_size
    b1, b2 = cfg.adam_beta1, cfg.adam_beta2
    lr_g, lr_c = cfg.lr_g(it), cfg.lr_c(it)
    vals = []
    self.counter_g += 1
    self.counter_v += 1
    vals.append(lr_g * math.sqrt(1.0 - b2 ** self.counter_g) / (1.0 - b1 ** self.counter_g))
    vals.append(cfg.value_lr_mul * lr_g * math.sqrt(1.0 - b2 ** self.counter_v) / (1.0 - b1 ** self.counter_v))
    for _ in range(I["citers"]):
      self.counter_c += 1
      vals.append(lr_c * math.sqrt(1.0 - b2 ** self.counter_c) / (1.0 - b1 ** self.counter_c))
    vals.append(float(it) / cfg.max_iter_step)         # progress
    if len(vals) <= 16:
      K.set_floats(I["hyper"], vals)                   # by value: no host buffer to keep alive, nothing to wait for
    else:
      for k, v in enumerate(vals):
        I["hyper"][k:k + 1].fill_(v)
    mem.stage_fresh()
    for k in range(I["citers"]):
      I["real"][k].copy_(mem.real_dataset.get_next_batch(B))
    self._itgraph.replay()
    return self._itout

  # ---- generator + value step (net.py:56-163, 222-239, 330) ---------------------------------
  def generator_forward(self, fake_input, states, noise, drop_f, drop_s, progress, is_train=1):
    c = self.policy.forward(fake_input, states, noise, drop_f, drop_s, is_train, progress, self.cfg)
    return c

  def generator_step(self, fake_input, states, noise, drop_f, drop_s, progress, lr_g, is_train=1, apply=True):
    """Returns dict(fake_output, new_states, g_loss, v_loss, ...) (device tensors, no sync).
    With enable_graphs() the returned tensors are the graph's static outputs: consume them
    before the next call."""
    if apply:
      self.counter_g += 1
      self.counter_v += 1
      self._set_lr("g", lr_g, self.counter_g)
      self._set_lr("v", self.cfg.value_lr_mul * lr_g, self.counter_v)
    graph = getattr(self, "_ggraph", None)
    if graph is not None and apply and is_train == 1 and fake_input.shape[0] == self._graph_B:
      gi = self._gi
      gi["img"].copy_(fake_input); gi["states"].copy_(states); gi["noise"].copy_(noise)
      gi["drop_f"].copy_(drop_f); gi["drop_s"].copy_(drop_s); gi["progress"].fill_(float(progress))
      graph.replay()
      if not self._graph_apply:
        self._apply("gv")
      return self._gout
    prog = torch.full((1,), float(progress), device=self.device)
    return self._generator_impl(fake_input, states, noise, drop_f, drop_s, prog, is_train, apply)

  def _generator_impl(self, *a):
    with K.weight_cache():                                       # padded first-layer weights: once per step and tensor
      return self._generator_sched(*a)

  def _critic_impl(self, *a):
    with K.weight_cache():
      return self._critic_sched(*a)

  def _generator_sched(self, fake_input, states, noise, drop_f, drop_s, progress, is_train, apply):
    # the two passes over the INPUT batch do not depend on the policy: parallel graph branches
    with K.fork(4):
      cc_in = self.critic.forward(fake_input)                    # fake_input_logit (stop_gradient) net.py:72-73
    with K.fork(5):
      v_old = self.value.forward(fake_input, states)             # old_value             net.py:79-84
    c = self.policy.forward(fake_input, states, noise, drop_f, drop_s, is_train, progress, self.cfg)
    with K.fork(5):                                              # the two passes over the OUTPUT batch: parallel
      v_new = self.value.forward(c.out, c.new_states)            # new_value             net.py:85-90
    cc_out = self.critic.forward(c.out)                          # fake_logit            net.py:70-71
    K.join()
    seeds, losses = K.rl_losses(cc_out.logit.view(-1), cc_in.logit.view(-1), v_old.logit.view(-1),
                                v_new.logit.view(-1), c.penalty, c.surrogate, c.new_states, self.cfg)
    # theta_v: v_loss = mean(advantage^2), advantage = stop_gradient(q) - old_value
    with K.fork(4):
      self.value.backward(v_old, seeds[2], param_grads=True)
    # theta_g: pathwise gradient through critic(fake_output) and value(fake_output, new_states):
    # two independent backward passes, parallel graph branches
    with K.fork(5):
      self.value.backward(v_new, seeds[1], param_grads=False)
      g_img_v = self.value.image_grad(v_new)
    self.critic.backward(cc_out, seeds[0], param_grads=False)
    g_img = self.critic.image_grad(cc_out)
    K.join()
    g_img = g_img + g_img_v
    self.policy.backward(c, g_img, seeds[4], seeds[3])
    if apply:
      self._apply("gv")
    return dict(fake_output=c.out, new_states=c.new_states, g_loss=losses[0], v_loss=losses[1], ctx=c,
                fake_logit=cc_out.logit, old_value=v_old.logit, new_value=v_new.logit, seeds=seeds)

  # ---- critic step (net.py:68-71, 151, 174-194, 245-251, 362) -------------------------------
  def critic_step(self, real, fake, alpha, lr_c, apply=True):
    """real, fake [B,64,64,3]; alpha [B] ~ U[0,1).  c_loss = mean(D(fake) - D(real)) + GP."""
    if apply:
      self.counter_c += 1
      self._set_lr("c", lr_c, self.counter_c)
    graph = getattr(self, "_cgraph", None)
    if graph is not None and apply and real.shape[0] == self._graph_B:
      ci = self._ci
      ci["real"].copy_(real); ci["fake"].copy_(fake); ci["alpha"].copy_(alpha)
      graph.replay()
      if not self._graph_apply:
        self._apply("c")
        self._ema_update(self._cout["c_average"])                # rank-local shard mean (a logging statistic)
      return self._cout
    return self._critic_impl(real, fake, alpha, apply)

  def _critic_sched(self, real, fake, alpha, apply):
    B = real.shape[0]
    lam = float(self.cfg.gradient_penalty_lambda)
    X = K.critic_inputs(real, fake, alpha)                         # real | fake | interpolated, one launch
    c = self.critic.forward(X)
    g_logit = self._g_logit_cache.get(B)                          # constant seeds: d c_loss / d logit
    if g_logit is None:
      g_logit = torch.empty(3 * B, device=real.device)
      g_logit[:B] = -1.0 / B
      g_logit[B:2 * B] = 1.0 / B
      g_logit[2 * B:] = 1.0
      self._g_logit_cache[B] = g_logit
    # the weight / bias gradients forked inside backward() keep running on their side streams while
    # the gradient-penalty chain (image gradient -> tangent pass) proceeds; the penalty's own wgrads
    # go to the same side stream per layer, so the accumulation order is fixed
    with K.deferred_join():
      self.critic.backward(c, g_logit, param_grads=True, sl=slice(0, 2 * B))
      sl = slice(2 * B, 3 * B)
      g = self.critic.image_grad(c, sl)                          # d inte_logit / d interpolated  net.py:181-183
      u, norm = K.gp_scale(g, lam)                               # d GP / d gradients             net.py:185-187
      if lam > 0:
        self.critic.gradient_penalty_grads(c, sl, u)
    logit = c.logit.view(-1)
    with K.fork(7):                                              # logging scalars: off the optimizer's path
      # emd (net.py:164), gradient penalty (185-187), c_loss (151, 194), c_average (165, forward of this step,
      # pre-update) and the moving average that is part of opt_c (166-168, 268-269): one launch
      sc = K.critic_scalars(logit, norm, lam, self.ema_state if apply else None)
      out = dict(emd=sc[0], gradient_penalty=sc[1], critic_gradient_norm=sc[2], c_loss=sc[3], c_average=sc[4], logits=logit)
    if apply:
      self._apply("c")
    K.join()
    return out

  # ---- GAN.train's inner loop (net.py:307-370): 1 generator+value step, cfg.citers critic steps
  def attach_memory(self, memory, generator=None):
    self.memory = memory
    self.rng = generator

  def train_iteration(self, it, giters=None, citers=None, lr_g=None):
    """One iteration of net.py:307-370 on the attached replay memory.  Returns device scalars
    (no host synchronisation inside)."""
    cfg = self.cfg
    B = cfg.batch_size
    progress = float(it) / cfg.max_iter_step
    if citers is None:
      citers = 100 if (it < cfg.critic_initialization or it % 500 == 0) else cfg.citers     # net.py:312-316
    if giters is None:
      giters = 100 if it == 0 else cfg.giters                                               # net.py:318-322
    if (getattr(self, "_itgraph", None) is not None and lr_g is None and it > 0 and giters == 1 and citers == self._it["citers"]):
      return self._iteration_replay(it)                                                      # the whole iteration: ONE graph
    lr_g = (0.0 if it == 0 else cfg.lr_g(it)) if lr_g is None else lr_g                      # net.py:327-328
    out = None
    for _ in range(giters):
      img, states, slots = self.memory.get_next_fake_batch(B)
      noise, drop_f, drop_s, _ = self.draw(B, self.rng)
      out = self.generator_step(img, states, noise, drop_f, drop_s, progress, lr_g)
      self.memory.replace_memory(out["fake_output"], out["new_states"], slots)              # net.py:340-342
    cout = None
    for _ in range(citers):
      fake, _ = self.memory.replay_fake_batch(B)                                            # replay_memory.py:159-173
      real = self.memory.real_dataset.get_next_batch(B)
      alpha = torch.rand(B, device=self.device, generator=self.rng)
      cout = self.critic_step(real, fake, alpha, cfg.lr_c(it))
    return dict(g_loss=out["g_loss"], v_loss=out["v_loss"], emd=cout["emd"] if cout else None,
                critic_gradient_norm=cout["critic_gradient_norm"] if cout else None,
                fake_output=out["fake_output"], new_states=out["new_states"])

  # ---- random draws the reference makes per step (explicit so tests can inject them) -------
  def draw(self, B, generator=None):
    dev = self.device
    noise = torch.rand(B, device=dev, generator=generator)                                     # z[:,0]  replay_memory.py:176-184
    keep = float(self.cfg.dropout_keep_prob)
    mk = lambda: torch.floor(keep + torch.rand(B, 4, 4, 256, device=dev, generator=generator)) / keep   # tf.nn.dropout
    alpha = torch.rand(B, device=dev, generator=generator)                                     # net.py:175-176
    return noise, mk(), mk(), alpha
