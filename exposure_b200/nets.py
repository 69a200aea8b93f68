"""Explicit (hand-scheduled) forward / backward of the Exposure networks on the C-ABI kernels.

  CriticNet  -- critics.py:42-98: global statistics -> enriched input -> 4 x [conv4x4s2 + lrelu]
                -> fc(128, lrelu) -> fc(1).  Used as WGAN critic (6 input channels) and, with the
                11 state channels, as value network (17 channels), exactly like cfg.critic /
                cfg.value = critic in config_example.py:98-99.
  PolicyNet  -- agent.py:41-125: shared filter feature extractor + 8 filter heads
                (filters.py:28-44), action-selection extractor + selector head, dropout always on.

No autograd: every backward signal is produced by an explicit call sequence (hard
part 2), so a train step is a fixed list of kernel launches (CUDA-graph capturable).  Variable
names and layouts follow the reference checkpoint so weights map 1:1."""
import math

import torch

from . import nn_ops as K
from . import ops as F

FEATURE_DIM = 4096            # cfg.feature_extractor_dims
FC1 = 128                     # cfg.fc1_size
CONV_CH = (32, 64, 128, 256)  # cfg.base_channels doubling (agent.py:24-33)
NUM_PARAMS = F.NUM_PARAMS
MASK_PARAMS = 6               # Filter.get_num_mask_parameters (filters.py:107-108)


class ParamStore:
  """All variables of one optimizer in ONE flat fp32 buffer (+ flat grads / Adam slots), with
  named views.  One NCCL all-reduce and one fused Adam launch per optimizer step."""

  def __init__(self, device):
    self.device = device
    self.specs = []          # (name, shape, fan_in, fan_out) ; fan_in == 0 -> zeros (biases)
    self.flat = None

  def add(self, name, shape, fan_in=0, fan_out=0):
    assert self.flat is None
    self.specs.append((name, tuple(shape), fan_in, fan_out))

  ALIGN = 64      # floats: 4 * world divides the padded size for every world <= 16 (exp_dp_allreduce_adam slices)

  def count(self):
    return sum(int(math.prod(s[1])) for s in self.specs)

  def padded(self):
    return (self.count() + self.ALIGN - 1) // self.ALIGN * self.ALIGN

  def finalize(self, seed=0, buffers=None):
    """buffers: (flat, grad, m, v) views of padded() floats each inside a shared ParamArena (several optimizers,
    one exchange); default: own buffers."""
    n = self.count()
    n_pad = self.padded()
    dev = self.device
    if buffers is None:
      buffers = tuple(torch.zeros(n_pad, device=dev) for _ in range(4))
    self.flat, self.grad, self.m, self.v = buffers
    assert all(b.numel() == n_pad for b in buffers)
    self.p, self.g, self.offsets = {}, {}, {}
    g = torch.Generator().manual_seed(seed)
    off = 0
    for name, shape, fi, fo in self.specs:
      k = int(math.prod(shape))
      self.p[name] = self.flat[off:off + k].view(shape)
      self.g[name] = self.grad[off:off + k].view(shape)
      self.offsets[name] = (off, k)
      if fi:   # tf.contrib.layers.xavier_initializer(uniform=True)
        lim = math.sqrt(6.0 / (fi + fo))
        self.p[name].copy_(((torch.rand(shape, generator=g) * 2 - 1) * lim).to(dev))
      off += k
    self.numel = n
    return self

  def zero_grad(self):
    self.grad.zero_()


def _add_conv_stack(store, scope, cin):
  names = []
  c = cin
  for i, co in enumerate(CONV_CH):
    base = "%s/Conv%s" % (scope, "" if i == 0 else "_%d" % i)
    store.add(base + "/weights", (4, 4, c, co), 16 * c, 16 * co)
    store.add(base + "/biases", (co,))
    names.append(base)
    c = co
  return names


class ParamArena:
  """Several ParamStores in ONE flat buffer (+ grads, Adam slots): the generator step trains theta_g and theta_v in
  the same sess.run (net.py:330-331), so both optimizers share one gradient exchange (7.35 M floats)."""

  def __init__(self, device, stores, seeds):
    sizes = [s.padded() for s in stores]
    total = sum(sizes)
    self.flat, self.grad, self.m, self.v = (torch.zeros(total, device=device) for _ in range(4))
    self.offsets, off = [], 0
    for s, n, seed in zip(stores, sizes, seeds):
      s.finalize(seed, buffers=tuple(b[off:off + n] for b in (self.flat, self.grad, self.m, self.v)))
      self.offsets.append(off)
      off += n
    self.numel = total


class _Ctx:
  pass


class ConvStack:
  """(x, vec) -> a1..a4 (post-lrelu activations) and everything needed to go back."""

  def __init__(self, store, scope, cin):
    self.store = store
    self.cin = cin
    self.names = _add_conv_stack(store, scope, cin)

  def W(self, i):
    return self.store.p[self.names[i] + "/weights"]

  def b(self, i):
    return self.store.p[self.names[i] + "/biases"]

  def forward(self, img, vec, drop_mul=None):
    """img [N,64,64,3], vec [N,cin-3].  Returns ctx with acts[0..3]; when drop_mul [N,4,4,256] is
    given ctx.feat = a4 * drop_mul (tf.nn.dropout, agent.py:36) else ctx.feat = a4 (flattened)."""
    c = _Ctx()
    c.img, c.vec = img, vec
    # ONE staging copy of the enriched first-layer input serves the forward and the weight gradient
    c.staged = K.stage_first_layer(img, vec, 0.5, CONV_CH[0])
    a = K.conv_fwd(img, self.W(0), self.b(0), vec=vec, shift=0.5, staged=c.staged)
    acts = [a]
    for i in (1, 2):
      a = K.conv_fwd(a, self.W(i), self.b(i))
      acts.append(a)
    if drop_mul is None:
      a = K.conv_fwd(a, self.W(3), self.b(3))
      c.feat = a.view(a.shape[0], -1)
    else:
      a, d = K.conv_fwd(a, self.W(3), self.b(3), post_mul=drop_mul)
      c.feat = d.view(d.shape[0], -1)
    acts.append(a)
    c.acts = acts
    return c

  def backward(self, c, delta4, param_grads=True, accumulate=False, sl=None, bias_tasks=None):
    """delta4 [N,4,4,256] = dL/d(pre-activation of conv 4).  Fills c.deltas[0..3].  The weight gradient of
    layer i only needs delta_i: it is forked onto a side stream and overlaps the rest of the dgrad chain
    (parallel branches of the CUDA graph).  The four bias gradients -- plus whatever column sums the caller
    hands in through `bias_tasks` (its FC layers) -- are ONE launch at the end (exp_colsum_multi)."""
    d = delta4
    deltas = [None, None, None, d]
    c.deltas = deltas
    for i in (3, 2, 1):
      if param_grads:
        with K.fork(i & 1):
          self._layer_wgrad(c, i, accumulate, sl)
      a_in = c.acts[i - 1]
      d = K.conv_dgrad(d, self.W(i), tuple(a_in.shape), a_in=a_in)
      deltas[i - 1] = d
    if param_grads:
      with K.fork(0):
        self._layer_wgrad(c, 0, accumulate, sl)
      with K.fork(1):
        K.colsum_multi(self._bias_tasks(c, accumulate, sl) + list(bias_tasks or []))
      K.join()

  def _layer_wgrad(self, c, i, accumulate, sl):
    s = sl if sl is not None else slice(None)
    g = self.store.g
    d = c.deltas[i][s]
    if i == 0:
      K.conv_wgrad(c.img[s], d, vec=c.vec[s], shift=0.5, out=g[self.names[0] + "/weights"], accumulate=accumulate,
                   staged=c.staged[s] if c.staged is not None else None)
    else:
      K.conv_wgrad(c.acts[i - 1][s], d, out=g[self.names[i] + "/weights"], accumulate=accumulate)

  def _bias_tasks(self, c, accumulate, sl):
    s = sl if sl is not None else slice(None)
    return [(c.deltas[i][s], self.store.g[self.names[i] + "/biases"], accumulate) for i in range(4)]

  def param_grads(self, c, accumulate=False, sl=None, bias_tasks=None):
    """wgrad + bias grads from the samples in `sl` (a slice over the batch; default all)."""
    for i in range(4):
      with K.fork(i & 1):
        self._layer_wgrad(c, i, accumulate, sl)
    with K.fork(1):
      K.colsum_multi(self._bias_tasks(c, accumulate, sl) + list(bias_tasks or []))
    K.join()

  def input_grad(self, c, sl=None):
    """dL/d(enriched, shifted layer-1 input) [n,64,64,cin] for the samples in sl."""
    s = sl if sl is not None else slice(None)
    d1 = c.deltas[0][s]
    n = d1.shape[0]
    return K.conv_dgrad(d1, self.W(0), (n, 64, 64, self.cin))

  def tangent(self, c, sl, t_img, t_vec):
    """Forward-mode tangents t1..t4 of the samples in sl for input tangent (t_img, t_vec)."""
    s = sl
    c.t_staged = K.stage_first_layer(t_img, t_vec, 0.0, CONV_CH[0])
    t = K.conv_fwd(t_img, self.W(0), None, vec=t_vec, shift=0.0, mask_ref=c.acts[0][s], staged=c.t_staged)
    ts = [t]
    for i in (1, 2, 3):
      t = K.conv_fwd(t, self.W(i), None, mask_ref=c.acts[i][s])
      ts.append(t)
    return ts

  def tangent_param_grads(self, c, sl, t_img, t_vec, ts):
    """Accumulate d<u, dD/dx>/dW_k = wgrad(t_{k-1}, delta_k)."""
    g = self.store.g
    with K.fork(0):
      K.conv_wgrad(t_img, c.deltas[0][sl], vec=t_vec, shift=0.0, out=g[self.names[0] + "/weights"], accumulate=True,
                   staged=getattr(c, "t_staged", None))
    for i in (1, 2, 3):
      with K.fork(i & 1):
        K.conv_wgrad(ts[i - 1], c.deltas[i][sl], out=g[self.names[i] + "/weights"], accumulate=True)
    K.join()


class CriticNet:
  """critics.py:42-98 (critic when n_states == 0, value network when n_states == 11)."""

  def __init__(self, store, scope, n_states=0):
    self.store = store
    self.n_states = n_states
    self.conv = ConvStack(store, scope, 3 + n_states + 3)
    self.fc1 = scope + "/fully_connected"
    self.fc2 = scope + "/fully_connected_1"
    store.add(self.fc1 + "/weights", (FEATURE_DIM, FC1), FEATURE_DIM, FC1)
    store.add(self.fc1 + "/biases", (FC1,))
    store.add(self.fc2 + "/weights", (FC1, 1), FC1, 1)
    store.add(self.fc2 + "/biases", (1,))

  def forward(self, images, states=None):
    p = self.store.p
    stats = K.stats_fwd(images)
    vec = stats if states is None else torch.cat([states, stats], dim=1)
    c = self.conv.forward(images, vec)
    c.stats = stats
    c.h = K.fc_fwd(c.feat, p[self.fc1 + "/weights"], p[self.fc1 + "/biases"], mode=K.FC_LRELU)
    c.logit = K.fc_fwd(c.h, p[self.fc2 + "/weights"], p[self.fc2 + "/biases"], mode=K.FC_LINEAR)
    return c

  def backward(self, c, g_logit, param_grads=True, accumulate=False, sl=None):
    """g_logit [N] = dL/dlogit.  Computes all deltas; parameter gradients from slice sl."""
    p = self.store.p
    c.d_fc2 = g_logit.reshape(-1, 1).contiguous()
    c.d_h = K.fc_dgrad(c.d_fc2, p[self.fc2 + "/weights"], mul_act=c.h)
    if param_grads:
      with K.fork(2):
        self._fc_wgrads(c, accumulate, sl)
    d4 = K.fc_dgrad(c.d_h, p[self.fc1 + "/weights"], mul_act=c.feat)
    self.conv.backward(c, d4.view(-1, 4, 4, CONV_CH[3]), param_grads=param_grads, accumulate=accumulate, sl=sl,
                       bias_tasks=self._fc_bias_tasks(c, accumulate, sl) if param_grads else None)
    K.join()

  def param_grads(self, c, accumulate=False, sl=None):
    with K.fork(2):
      self._fc_wgrads(c, accumulate, sl)
    self.conv.param_grads(c, accumulate=accumulate, sl=sl, bias_tasks=self._fc_bias_tasks(c, accumulate, sl))
    K.join()

  def _fc_wgrads(self, c, accumulate, sl):
    s = sl if sl is not None else slice(None)
    g = self.store.g
    K.fc_wgrad(c.feat[s], c.d_h[s], out=g[self.fc1 + "/weights"], accumulate=accumulate)
    K.fc_wgrad(c.h[s], c.d_fc2[s], out=g[self.fc2 + "/weights"], accumulate=accumulate)

  def _fc_bias_tasks(self, c, accumulate, sl):
    s = sl if sl is not None else slice(None)
    g = self.store.g
    return [(c.d_h[s], g[self.fc1 + "/biases"], accumulate), (c.d_fc2[s], g[self.fc2 + "/biases"], accumulate)]

  def image_grad(self, c, sl=None, g_direct_extra=None):
    """dL/dimages for the samples in sl: layer-1 dgrad, image channels + J_stats^T of the
    three statistic channels (tf.gradients through critics.py:48-87)."""
    s = sl if sl is not None else slice(None)
    if K.first_layer_split(self.conv.cin):
      # the image channels' gradient + the per-image pixel sums of the tiled channels' (exp_conv_first_dgrad); the three
      # statistic channels go back through J_stats^T (exp_stats_bwd), the state channels have no producer
      d1 = c.deltas[0][s]
      g_img, g_vec = K.conv_first_dgrad(d1, self.conv.W(0), self.conv.cin - 3, (64, 64))
      return K.stats_bwd(c.img[s], c.stats[s], g_vec[:, self.conv.cin - 6:].contiguous(), g_direct=g_img)
    g_in = self.conv.input_grad(c, sl)                         # [n,64,64,cin]
    # image channels pass through, the three tiled statistic channels are summed per image and pulled back through
    # J_stats^T -- one launch (exp_stats_bwd_gin)
    return K.stats_bwd_gin(c.img[s], c.stats[s], g_in)

  def gradient_penalty_grads(self, c, sl, u):
    """Accumulate d<u, d logit/d image>/dtheta for the samples in sl (net.py:181-194):
    forward-mode tangent pass, then wgrad(tangent_{k-1}, delta_k)."""
    p, g = self.store.p, self.store.g
    img, stats = c.img[sl], c.stats[sl]
    dstat = K.stats_jvp(img, stats, u)
    if self.n_states:
      dstat = torch.cat([torch.zeros(u.shape[0], self.n_states, device=u.device), dstat], dim=1)
    ts = self.conv.tangent(c, sl, u, dstat)
    t4 = ts[3].view(ts[3].shape[0], -1)
    th = K.fc_fwd(t4, p[self.fc1 + "/weights"], None, mode=K.FC_TANGENT, mask_ref=c.h[sl])
    c.gp_keep = (ts, t4, th, dstat, u)       # read by forked blocks: alive until the caller's join
    with K.fork(2):
      K.fc_wgrad(t4, c.d_h[sl], out=g[self.fc1 + "/weights"], accumulate=True)
      K.fc_wgrad(th, c.d_fc2[sl], out=g[self.fc2 + "/weights"], accumulate=True)
    self.conv.tangent_param_grads(c, sl, u, dstat, ts)
    K.join()


class PolicyNet:
  """agent.py:41-125 with cfg.shared_feature_extractor = True."""

  def __init__(self, store, n_states=11, scope="generator"):
    self.store = store
    self.n_filters = len(NUM_PARAMS)
    self.fe = ConvStack(store, scope, 3 + n_states)
    self.se = ConvStack(store, scope + "/action_selection", 3 + n_states)
    # the 8 fc1 layers of the filter heads are stored as ONE [4096, 8*128] matrix (column block j
    # == generator/filter_j/fc1/weights) so that they run as a single GEMM
    self.fc1_all = scope + "/filter_fc1_all"
    store.add(self.fc1_all + "/weights", (FEATURE_DIM, self.n_filters * FC1), FEATURE_DIM, FC1)
    store.add(self.fc1_all + "/biases", (self.n_filters * FC1,))
    self.fc2 = []
    self.out_dims = [n + MASK_PARAMS for n in NUM_PARAMS]
    for j, od in enumerate(self.out_dims):
      name = "%s/filter_%d/fc2" % (scope, j)
      store.add(name + "/weights", (FC1, od), FC1, od)
      store.add(name + "/biases", (od,))
      self.fc2.append(name)
    self.sfc1 = scope + "/action_selection/selector_fc1"
    self.sfc2 = scope + "/action_selection/selector_fc2"
    store.add(self.sfc1 + "/weights", (FEATURE_DIM, FC1), FEATURE_DIM, FC1)
    store.add(self.sfc1 + "/biases", (FC1,))
    store.add(self.sfc2 + "/weights", (FC1, self.n_filters), FC1, self.n_filters)
    store.add(self.sfc2 + "/biases", (self.n_filters,))
    self.ostride = max(self.out_dims)          # 30
    self._heads = None                         # K.HeadsLayout, built on first use (needs the finalized store)

  def heads(self):
    if self._heads is None:
      self._heads = K.HeadsLayout(self.store, self.fc2, self.out_dims, list(NUM_PARAMS), FC1, MASK_PARAMS)
    return self._heads

  def forward(self, img, states, noise, drop_f, drop_s, is_train, progress, cfg, high_res=None):
    """img [B,64,64,3], states [B,11], noise [B] (= z[:,0]), drop_* [B,4,4,256] in {0,2}.
    Returns ctx with .out (filtered image), .new_states, .surrogate, .penalty, .ids, .pdf, ..."""
    p = self.store.p
    B = img.shape[0]
    c = _Ctx()
    c.img, c.states, c.progress, c.cfg = img, states, progress, cfg
    c.drop_f, c.drop_s = drop_f, drop_s
    # the action-selection tower (agent.py:80-99) has its own weights and only meets the filter-head
    # tower at the policy head: a parallel branch
    with K.fork(6):
      c.s = self.se.forward(img, states, drop_mul=drop_s)
      c.hs = K.fc_fwd(c.s.feat, p[self.sfc1 + "/weights"], p[self.sfc1 + "/biases"], mode=K.FC_LRELU)
      c.sel_logits = K.fc_fwd(c.hs, p[self.sfc2 + "/weights"], p[self.sfc2 + "/biases"], mode=K.FC_LINEAR)
    c.f = self.fe.forward(img, states, drop_mul=drop_f)
    # filter heads (filters.py:28-44)
    c.H = K.fc_fwd(c.f.feat, p[self.fc1_all + "/weights"], p[self.fc1_all + "/biases"], mode=K.FC_LRELU)
    c.O = K.heads_fc2_fwd(self.heads(), c.H)          # all 8 fc2 layers: [B, 8, 30], one launch
    K.join()
    # action selection (agent.py:100-122)
    c.pdf, c.ids, c.surrogate, c.entropy, c.pen_head, c.new_states = K.policy_head_fwd(
        c.sel_logits, noise, states, is_train, progress, cfg)
    # only the selected filter is evaluated (agent.py:124-125 computes all 8 and one-hot sums); an id of -1
    # (pdf_sample with u == 0: all-zero one-hot row) makes the step kernels write a black image themselves
    c.masking = bool(getattr(cfg, "masking", False))
    c.logits_sel, c.mask_logits_sel = K.heads_select(self.heads(), c.O, c.ids, F.PSTRIDE, c.masking)
    c.params = F.filter_regress_fwd(c.logits_sel, c.ids)
    if not c.masking:
      c.out = F.filter_fwd(img, c.params, c.ids)
      if high_res is not None:
        c.high_res_out = F.filter_fwd(high_res, c.params, c.ids)
    else:
      # cfg.masking (filters.py:62-148): the 6 mask logits of the selected filter are its fc2
      # outputs [n, n+6) (c.mask_logits_sel); the mask is evaluated inside the masked step kernel
      c.mask_cfg = (float(cfg.maximum_sharpness), float(cfg.minimum_strength))
      c.out = F.filter_masked_fwd(img, c.params, c.mask_logits_sel, c.ids, *c.mask_cfg, True)
      if high_res is not None:
        c.high_res_out = F.filter_masked_fwd(high_res, c.params, c.mask_logits_sel, c.ids, *c.mask_cfg, True)
    c.pen_img = K.overexposure_fwd(c.out)
    c.penalty = c.pen_img + c.pen_head
    return c

  def backward(self, c, g_out, g_surrogate, g_penalty):
    """g_out [B,64,64,3] = dL/d filtered image EXCLUDING the over-exposure penalty path (added
    here), g_surrogate / g_penalty [B].  Fills the generator gradients (overwrite)."""
    p, g = self.store.p, self.store.g
    B = c.img.shape[0]
    with K.fork(3):                      # the selector tower is independent of the filter-head tower
      self._selector_backward(c, g_surrogate, g_penalty)
    g_img = K.overexposure_bwd(c.out, g_penalty, g_in=g_out)
    g_mask = None
    if not c.masking:
      _, g_params = F.filter_bwd(c.img, g_img, c.params, c.ids, need_gx=False)
    else:
      _, g_params, g_mask = F.filter_masked_bwd(c.img, g_img, c.params, c.mask_logits_sel, c.ids, *c.mask_cfg, True,
                                                need_gx=False)
      g_mask = g_mask[:, :MASK_PARAMS].contiguous()
    g_logits_sel = F.filter_regress_bwd(c.logits_sel, g_params, c.ids)           # [B,24], entries >= n are 0
    # backward of the select + the 8 fc2 layers in one launch: only the head an image selected receives its
    # gradient (an id of -1 none), dH comes back with lrelu' applied, fc2 weight / bias gradients are overwritten
    d_H = K.heads_fc2_bwd(self.heads(), c.H, c.ids, g_logits_sel, g_mask)
    with K.fork(2):
      K.fc_wgrad(c.f.feat, d_H, out=g[self.fc1_all + "/weights"])
    a4f = c.f.acts[3].view(B, -1)
    d4 = K.fc_dgrad(d_H, p[self.fc1_all + "/weights"], mul_act=a4f, mul_plain=c.drop_f.view(B, -1))
    self.fe.backward(c.f, d4.view(B, 4, 4, CONV_CH[3]), bias_tasks=[(d_H, g[self.fc1_all + "/biases"], False)])
    K.join()

  def _selector_backward(self, c, g_surrogate, g_penalty):
    p, g = self.store.p, self.store.g
    B = c.img.shape[0]
    g_sel = K.policy_head_bwd(c.sel_logits, c.ids, g_surrogate, g_penalty, c.progress, c.cfg)
    K.fc_wgrad(c.hs, g_sel, out=g[self.sfc2 + "/weights"])
    d_hs = K.fc_dgrad(g_sel, p[self.sfc2 + "/weights"], mul_act=c.hs)
    K.fc_wgrad(c.s.feat, d_hs, out=g[self.sfc1 + "/weights"])
    a4s = c.s.acts[3].view(B, -1)
    d4s = K.fc_dgrad(d_hs, p[self.sfc1 + "/weights"], mul_act=a4s, mul_plain=c.drop_s.view(B, -1))
    self.se.backward(c.s, d4s.view(B, 4, 4, CONV_CH[3]),
                     bias_tasks=[(g_sel, g[self.sfc2 + "/biases"], False), (d_hs, g[self.sfc1 + "/biases"], False)])
