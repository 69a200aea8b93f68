"""torch-tensor front end of the C ABI (device pointers, current stream, autograd glue).

PyTorch here is plumbing only: it owns device memory and streams.  Every function enqueues
exactly the CUDA kernels of exposure_b200/csrc on torch's current stream."""
import torch

from . import _cabi
from ._cabi import MAX_FILTER_PARAMS, VARIANT_AUTO

NUM_PARAMS = (1, 1, 3, 1, 8, 1, 1, 24)      # cfg.filters order (config_example.py:22-25)
NUM_PARAMS_ALL = NUM_PARAMS + (2, 1)        # + LevelFilter (id 8), VignetFilter (id 9)
PSTRIDE = MAX_FILTER_PARAMS
MASK_PARAMS = 6                             # Filter.get_num_mask_parameters (filters.py:107-108)

# running count of kernels launched through this module (bench.py reports it)
launch_count = 0
# when set to a list, every filter-step launch appends (name, algorithmic_bytes, ev0, ev1)
# with CUDA events recorded on the launching stream (bench.py's live roofline measurement)
event_log = None
FILTER_NAMES = ("exposure", "gamma", "wb", "satplus", "tone", "contrast", "wnb", "color", "level", "vignet")


class _Timed:
  """Brackets one launch with CUDA events on the current stream when event_log is active."""

  def __init__(self, kind, ids, nbytes):
    self.on = event_log is not None
    if self.on:
      self.name = "%s_%s" % (kind, ids if isinstance(ids, str) else
                             (FILTER_NAMES[ids] if isinstance(ids, int) else "select"))
      self.nbytes = nbytes
      self.e0 = torch.cuda.Event(enable_timing=True)
      self.e1 = torch.cuda.Event(enable_timing=True)

  def __enter__(self):
    if self.on:
      self.e0.record()

  def __exit__(self, *a):
    if self.on:
      self.e1.record()
      event_log.append((self.name, self.nbytes, self.e0, self.e1))


def _stream():
  return torch.cuda.current_stream().cuda_stream


def _chk_img(t, name):
  if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 4 and t.shape[3] == 3 and t.is_contiguous()):
    raise ValueError("%s must be a contiguous CUDA float32 NHWC [B,H,W,3] tensor, got %s %s %s" %
                     (name, tuple(t.shape), t.dtype, t.device))


def _ids_arg(ids, B):
  """ids: int (uniform filter id) or CUDA int32 tensor [B] -> (device ptr or None, uniform id)."""
  if isinstance(ids, int):
    return None, ids
  if not (ids.is_cuda and ids.dtype == torch.int32 and ids.shape == (B,) and ids.is_contiguous()):
    raise ValueError("ids must be an int or a contiguous CUDA int32 tensor [B]")
  return ids.data_ptr(), 0


def _chk_mat(t, B, name):
  if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.shape[0] == B and t.stride(1) == 1):
    raise ValueError("%s must be CUDA float32 [B, n] with unit inner stride" % name)
  return t.stride(0)


def filter_regress_fwd(logits, ids):
  """filter_param_regressor: logits [B, >=n] -> params [B, 24] (entries >= n are 0)."""
  global launch_count
  B = logits.shape[0]
  ls = _chk_mat(logits, B, "logits")
  params = torch.zeros(B, PSTRIDE, device=logits.device, dtype=torch.float32)
  idp, uid = _ids_arg(ids, B)
  _cabi.check(_cabi.lib().exp_filter_regress_fwd(logits.data_ptr(), ls, params.data_ptr(), PSTRIDE, idp, uid, B,
                                                 _stream()), "exp_filter_regress_fwd")
  launch_count += 1
  return params


def filter_regress_bwd(logits, gparams, ids):
  global launch_count
  B = logits.shape[0]
  ls = _chk_mat(logits, B, "logits")
  ps = _chk_mat(gparams, B, "gparams")
  if not logits.is_contiguous():
    raise ValueError("logits must be contiguous")
  glogits = torch.empty_like(logits)
  idp, uid = _ids_arg(ids, B)
  _cabi.check(_cabi.lib().exp_filter_regress_bwd(logits.data_ptr(), ls, gparams.data_ptr(), ps, glogits.data_ptr(),
                                                 idp, uid, B, _stream()), "exp_filter_regress_bwd")
  launch_count += 1
  return glogits


OPT_LOGITS = 0x100     # EXP_OPT_LOGITS: params are raw regressor logits (regressor fused in-kernel)


def filter_fwd(x, params, ids, out=None, variant=VARIANT_AUTO, logits=False):
  """y = process_{ids}(x, params).  x: [B,H,W,3]; params: [B, >=n]; ids: int or int32 [B].
  logits=True: `params` are the raw regressor logits (filter_param_regressor runs in-kernel)."""
  global launch_count
  _chk_img(x, "x")
  B, H, W, _ = x.shape
  ps = _chk_mat(params, B, "params")
  y = torch.empty_like(x) if out is None else out
  idp, uid = _ids_arg(ids, B)
  with _Timed("filter_fwd", ids, B * H * W * 24):
    _cabi.check(_cabi.lib().exp_filter_fwd(x.data_ptr(), y.data_ptr(), params.data_ptr(), ps, idp, uid, B, H, W,
                                           variant | (OPT_LOGITS if logits else 0), _stream()), "exp_filter_fwd")
  launch_count += 1
  return y


def filter_chain_fwd(x, params, ids, out=None, logits=False):
  """S filter steps in ONE pass: x [B,H,W,3]; params [S,B,24]; ids int32 [S,B]."""
  global launch_count
  _chk_img(x, "x")
  B, H, W, _ = x.shape
  S = ids.shape[0]
  if not (params.is_cuda and params.dtype == torch.float32 and params.is_contiguous() and tuple(params.shape) == (S, B, PSTRIDE)):
    raise ValueError("params must be a contiguous CUDA float32 [S,B,%d] tensor" % PSTRIDE)
  if not (ids.is_cuda and ids.dtype == torch.int32 and ids.is_contiguous() and tuple(ids.shape) == (S, B)):
    raise ValueError("ids must be a contiguous CUDA int32 [S,B] tensor")
  y = torch.empty_like(x) if out is None else out
  with _Timed("filter_chain_fwd", "fused%d" % S, B * H * W * 24):
    _cabi.check(_cabi.lib().exp_filter_chain_fwd(x.data_ptr(), y.data_ptr(), params.data_ptr(), PSTRIDE, ids.data_ptr(),
                                                 S, B, H, W, OPT_LOGITS if logits else 0, _stream()),
                "exp_filter_chain_fwd")
  launch_count += 1
  return y


_workspaces = {}


def _workspace(dev, nbytes):
  """Zero-initialised, self-cleaning workspace per (device, stream)."""
  key = (dev, torch.cuda.current_stream().cuda_stream)
  ws = _workspaces.get(key)
  if ws is None or ws.numel() < nbytes:
    ws = torch.zeros(max(nbytes, 1 << 16), dtype=torch.uint8, device=dev)
    _workspaces[key] = ws
  return ws


OPT_NO_STATIC_CHAIN = 0x200   # EXP_OPT_NO_STATIC_CHAIN: uniform chains never take a compile-time instantiation (tests)


def filter_chain_fwd_bwd(x, gy, params, ids, need_y=True, need_gx=True, logits=False, y_out=None, gx_out=None,
                         gparams_out=None, static_chain=True):
  """Whole chain forward + backward in ONE pass over the pixels: x, gy [B,H,W,3]; params [S,B,24];
  ids: CUDA int32 [S,B] (per-image filter ids, exp_filter_chain_fwd_bwd) or a list / tuple of S ints (the same
  sequence for every image, exp_filter_chain_fwd_bwd_uniform: sequences with a compile-time instantiation run
  the specialised kernel unless static_chain=False).  Returns (y or None, gx or None, gparams [S,B,24])."""
  global launch_count
  _chk_img(x, "x")
  _chk_img(gy, "gy")
  B, H, W, _ = x.shape
  uniform = isinstance(ids, (list, tuple))
  S = len(ids) if uniform else ids.shape[0]
  if not (params.is_cuda and params.dtype == torch.float32 and params.is_contiguous() and tuple(params.shape) == (S, B, PSTRIDE)):
    raise ValueError("params must be a contiguous CUDA float32 [S,B,%d] tensor" % PSTRIDE)
  if not uniform and not (ids.is_cuda and ids.dtype == torch.int32 and ids.is_contiguous() and tuple(ids.shape) == (S, B)):
    raise ValueError("ids must be a contiguous CUDA int32 [S,B] tensor or a list of S ints")
  y = (torch.empty_like(x) if y_out is None else y_out) if need_y else None
  gx = (torch.empty_like(x) if gx_out is None else gx_out) if need_gx else None
  gparams = torch.zeros(S, B, PSTRIDE, device=x.device, dtype=torch.float32) if gparams_out is None else gparams_out
  l = _cabi.lib()
  ws = _workspace(x.device, l.exp_filter_chain_fwd_bwd_workspace_bytes(S, B, H, W))
  opts = OPT_LOGITS if logits else 0
  with _Timed("filter_chain_fwd_bwd", "fused%d" % S, B * H * W * (24 + (12 if need_y else 0) + (12 if need_gx else 0))):
    if uniform:
      import ctypes
      ids_host = (ctypes.c_int * S)(*[int(i) for i in ids])
      _cabi.check(l.exp_filter_chain_fwd_bwd_uniform(
          x.data_ptr(), gy.data_ptr(), y.data_ptr() if need_y else None, gx.data_ptr() if need_gx else None,
          params.data_ptr(), PSTRIDE, ctypes.cast(ids_host, ctypes.c_void_p), S, B, H, W, gparams.data_ptr(), ws.data_ptr(),
          ws.numel(), opts | (0 if static_chain else OPT_NO_STATIC_CHAIN), _stream()), "exp_filter_chain_fwd_bwd_uniform")
    else:
      _cabi.check(l.exp_filter_chain_fwd_bwd(x.data_ptr(), gy.data_ptr(), y.data_ptr() if need_y else None,
                                             gx.data_ptr() if need_gx else None, params.data_ptr(), PSTRIDE, ids.data_ptr(),
                                             S, B, H, W, gparams.data_ptr(), ws.data_ptr(), ws.numel(), opts, _stream()),
                  "exp_filter_chain_fwd_bwd")
  launch_count += 1
  return y, gx, gparams


def filter_bwd(x, gy, params, ids, need_gx=True, gx_out=None, variant=VARIANT_AUTO, logits=False, gparams_out=None):
  """Returns (gx or None, gparams [B,24]).  Deterministic parameter-gradient reduction.
  logits=True: `params` are raw regressor logits and the returned gparams are dL/dlogits."""
  global launch_count
  _chk_img(x, "x")
  _chk_img(gy, "gy")
  B, H, W, _ = x.shape
  ps = _chk_mat(params, B, "params")
  gx = (torch.empty_like(x) if gx_out is None else gx_out) if need_gx else None
  # the ABI uses ONE row stride for params and gparams
  if gparams_out is None:
    gparams = torch.zeros(B, ps, device=x.device, dtype=torch.float32)
  else:
    gparams = gparams_out
    if _chk_mat(gparams, B, "gparams_out") != ps:
      raise ValueError("gparams_out must have the same row stride as params (%d)" % ps)
  l = _cabi.lib()
  nbytes = l.exp_filter_bwd_workspace_bytes(B, H, W)
  ws = _workspace(x.device, nbytes)
  idp, uid = _ids_arg(ids, B)
  with _Timed("filter_bwd" if need_gx else "filter_bwd_paramonly", ids, B * H * W * (36 if need_gx else 24)):
    _cabi.check(l.exp_filter_bwd(x.data_ptr(), gy.data_ptr(), gx.data_ptr() if need_gx else None,
                                 gparams.data_ptr(), params.data_ptr(), ps, idp, uid, B, H, W, ws.data_ptr(),
                                 ws.numel(), variant | (OPT_LOGITS if logits else 0), _stream()), "exp_filter_bwd")
  launch_count += 1
  return gx, gparams


def _chk_mask_logits(mask_logits, B):
  if mask_logits is None:
    return None, MASK_PARAMS
  return mask_logits.data_ptr(), _chk_mat(mask_logits, B, "mask_logits")


def filter_masked_fwd(x, params, mask_logits, ids, max_sharpness=1.0, min_strength=0.3, masking=True, out=None,
                      want_mask=False, logits=False):
  """Filter.apply with cfg.masking (filters.py:62-99): y = lerp(x, process(x, params), get_mask(x, mask_logits)).
  mask_logits: raw fc2 outputs [B, >=6] (None = zeros).  Returns y, or (y, mask [B,H,W,1]) with want_mask."""
  global launch_count
  _chk_img(x, "x")
  B, H, W, _ = x.shape
  ps = _chk_mat(params, B, "params")
  mp, ms = _chk_mask_logits(mask_logits, B)
  y = torch.empty_like(x) if out is None else out
  mask = torch.empty(B, H, W, 1, device=x.device, dtype=torch.float32) if want_mask else None
  idp, uid = _ids_arg(ids, B)
  with _Timed("filter_masked_fwd", ids, B * H * W * 24):
    _cabi.check(_cabi.lib().exp_filter_masked_fwd(
        x.data_ptr(), y.data_ptr(), mask.data_ptr() if want_mask else None, params.data_ptr(), ps, mp, ms, idp, uid,
        B, H, W, float(max_sharpness), float(min_strength), int(bool(masking)), OPT_LOGITS if logits else 0, _stream()),
                "exp_filter_masked_fwd")
  launch_count += 1
  return (y, mask) if want_mask else y


def filter_mask(x, mask_logits, ids, max_sharpness=1.0, min_strength=0.3, masking=True):
  """Filter.get_mask alone (filters.py:110-148 / 354-396): [B,H,W,1]."""
  global launch_count
  _chk_img(x, "x")
  B, H, W, _ = x.shape
  mp, ms = _chk_mask_logits(mask_logits, B)
  mask = torch.empty(B, H, W, 1, device=x.device, dtype=torch.float32)
  idp, uid = _ids_arg(ids, B)
  _cabi.check(_cabi.lib().exp_filter_masked_fwd(
      x.data_ptr(), None, mask.data_ptr(), None, PSTRIDE, mp, ms, idp, uid, B, H, W, float(max_sharpness),
      float(min_strength), int(bool(masking)), 0, _stream()), "exp_filter_masked_fwd[mask]")
  launch_count += 1
  return mask


def filter_masked_bwd(x, gy, params, mask_logits, ids, max_sharpness=1.0, min_strength=0.3, masking=True,
                      need_gx=True, logits=False):
  """Returns (gx or None, gparams [B, pstride], gmask_logits [B, mstride])."""
  global launch_count
  _chk_img(x, "x")
  _chk_img(gy, "gy")
  B, H, W, _ = x.shape
  ps = _chk_mat(params, B, "params")
  mp, ms = _chk_mask_logits(mask_logits, B)
  gx = torch.empty_like(x) if need_gx else None
  gparams = torch.zeros(B, ps, device=x.device, dtype=torch.float32)
  gmask = torch.zeros(B, ms, device=x.device, dtype=torch.float32)
  l = _cabi.lib()
  ws = _workspace(x.device, l.exp_filter_bwd_workspace_bytes(B, H, W))
  idp, uid = _ids_arg(ids, B)
  with _Timed("filter_masked_bwd" if need_gx else "filter_masked_bwd_paramonly", ids, B * H * W * (36 if need_gx else 24)):
    _cabi.check(l.exp_filter_masked_bwd(
        x.data_ptr(), gy.data_ptr(), gx.data_ptr() if need_gx else None, gparams.data_ptr(), gmask.data_ptr(),
        params.data_ptr(), ps, mp, ms, idp, uid, B, H, W, float(max_sharpness), float(min_strength),
        int(bool(masking)), ws.data_ptr(), ws.numel(), OPT_LOGITS if logits else 0, _stream()), "exp_filter_masked_bwd")
  launch_count += 1
  return gx, gparams, gmask


class FilterMaskedFn(torch.autograd.Function):
  """autograd node for one masked filter step: (x, params, mask_logits) -> y."""

  @staticmethod
  def forward(ctx, x, params, mask_logits, ids, max_sharpness, min_strength):
    ctx.ids, ctx.cfgv = ids, (max_sharpness, min_strength)
    x, mask_logits = x.contiguous(), mask_logits.contiguous()
    ctx.save_for_backward(x, params, mask_logits)
    return filter_masked_fwd(x, params, mask_logits, ids, max_sharpness, min_strength, True)

  @staticmethod
  def backward(ctx, gy):
    x, params, mask_logits = ctx.saved_tensors
    gx, gparams, gmask = filter_masked_bwd(x, gy.contiguous(), params, mask_logits, ctx.ids, ctx.cfgv[0], ctx.cfgv[1],
                                           True, need_gx=ctx.needs_input_grad[0])
    return gx, gparams, gmask, None, None, None


class FilterProcessFn(torch.autograd.Function):
  """autograd node for one filter step: (x, params) -> y, backward through exp_filter_bwd."""

  @staticmethod
  def forward(ctx, x, params, ids):
    ctx.ids = ids
    ctx.save_for_backward(x, params)
    return filter_fwd(x.contiguous(), params, ids)

  @staticmethod
  def backward(ctx, gy):
    x, params = ctx.saved_tensors
    need_gx = ctx.needs_input_grad[0]
    gx, gparams = filter_bwd(x.contiguous(), gy.contiguous(), params, ctx.ids, need_gx=need_gx)
    return gx, gparams[:, :params.shape[1]] if params.shape[1] != PSTRIDE else gparams, None


class FilterChainFn(torch.autograd.Function):
  """autograd node for a whole chain of S filter steps: (x, logits [S,B,24]) -> x_S with ids [S,B].
  forward: exp_filter_chain_fwd (24 B/pixel for all S steps); backward: exp_filter_chain_fwd_bwd with the
  output store skipped (36 B/pixel) -- 60 B/pixel per training step for the whole chain, no intermediate
  image is ever saved for the backward (only x)."""

  @staticmethod
  def forward(ctx, x, logits, ids):
    x, logits = x.contiguous(), logits.contiguous()
    ctx.ids = ids
    ctx.save_for_backward(x, logits)
    return filter_chain_fwd(x, logits, ids, logits=True)

  @staticmethod
  def backward(ctx, gy):
    x, logits = ctx.saved_tensors
    _, gx, glogits = filter_chain_fwd_bwd(x, gy.contiguous(), logits, ctx.ids, need_y=False,
                                          need_gx=ctx.needs_input_grad[0], logits=True)
    return gx, glogits, None


def filter_chain(x, logits, ids):
  """Differentiable S-step filter chain (torch autograd): x [B,H,W,3], logits [S,B,24] raw regressor inputs,
  ids int32 [S,B]."""
  return FilterChainFn.apply(x, logits, ids)


class FilterRegressFn(torch.autograd.Function):
  """autograd node for filter_param_regressor: logits -> params[B,24]."""

  @staticmethod
  def forward(ctx, logits, ids):
    ctx.ids = ids
    logits = logits.contiguous()
    ctx.save_for_backward(logits)
    return filter_regress_fwd(logits, ids)

  @staticmethod
  def backward(ctx, gparams):
    (logits,) = ctx.saved_tensors
    return filter_regress_bwd(logits, gparams.contiguous(), ctx.ids), None
