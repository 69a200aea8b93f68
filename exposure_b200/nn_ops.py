"""torch-tensor front end of the network / head primitives of the C ABI.

Plumbing only: device pointers, current stream, output allocation.  See
include/exposure_b200.h for the semantics of every call.  2-D operands may be column
slices of a wider matrix (unit inner stride); their row stride is passed as the leading
dimension."""
import contextlib
import os

import torch

from . import _cabi
from . import ops as _ops

_ws = {}


def _stream():
  return torch.cuda.current_stream().cuda_stream


def _workspace(dev, nbytes):
  key = (dev, torch.cuda.current_stream().cuda_stream)
  ws = _ws.get(key)
  if ws is None or ws.numel() < nbytes:
    ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=dev)
    _ws[key] = ws
  return ws


_scratch_bufs = {}


def _scratch(dev, tag, numel):
  """Persistent float32 scratch per (device, stream, tag): staging copies that must outlive one call
  (safe inside CUDA-graph capture once the eager warm-up has sized them)."""
  key = (dev, torch.cuda.current_stream().cuda_stream, tag)
  buf = _scratch_bufs.get(key)
  if buf is None or buf.numel() < numel:
    buf = torch.empty(int(numel), dtype=torch.float32, device=dev)
    _scratch_bufs[key] = buf
  return buf


def _p(t):
  return None if t is None else t.data_ptr()


# ---- parallel branches ------------------------------------------------------------------
# `with fork(i): ...` runs the block on side stream i, ordered after everything queued so far on the
# current stream; `join()` makes the current stream wait for the blocks it forked.  Under CUDA-graph
# capture this turns independent work (weight / bias gradients vs. the dgrad chain, the two towers of
# the policy network) into parallel graph branches, which matters because most kernels of the train
# step are single-wave and 15-20 us long.  Discipline (keeps the caching allocator safe without
# record_stream): a forked block only reads tensors that stay alive until the join, and every block
# starts with wait_stream(parent).  EXPOSURE_FORK=0 serialises everything (A/B switch).
FORK_ENABLED = os.environ.get("EXPOSURE_FORK", "1") != "0"
_side_streams = {}
_pending = {}


@contextlib.contextmanager
def fork(idx=0):
  if not FORK_ENABLED:
    yield
    return
  parent = torch.cuda.current_stream()
  key = (parent.device, idx)
  side = _side_streams.get(key)
  if side is None:
    side = _side_streams[key] = torch.cuda.Stream(device=parent.device)
  if side.cuda_stream == parent.cuda_stream:          # already on that side stream: run inline
    yield
    return
  side.wait_stream(parent)
  with torch.cuda.stream(side):
    yield
  _pending.setdefault(parent.cuda_stream, []).append(side)


_deferred = set()


def join():
  cur = torch.cuda.current_stream()
  if cur.cuda_stream in _deferred:
    return
  for side in _pending.pop(cur.cuda_stream, []):
    cur.wait_stream(side)


@contextlib.contextmanager
def deferred_join():
  """Inside the block, join() on the CURRENT stream is postponed to the end of the block, so forked
  work keeps overlapping whatever the current stream does next.  Only safe when later forks that touch
  the same outputs go to the same side stream (nets.py sends layer i to stream i & 1 every time) and
  every tensor a forked block reads stays alive until the block ends."""
  cur = torch.cuda.current_stream()
  nested = cur.cuda_stream in _deferred
  _deferred.add(cur.cuda_stream)
  try:
    yield
  finally:
    if not nested:
      _deferred.discard(cur.cuda_stream)
      join()


def _chk(t, name, dims=None):
  if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
    raise ValueError("%s must be a contiguous CUDA float32 tensor" % name)
  if dims is not None and t.dim() != dims:
    raise ValueError("%s must have %d dims, got %s" % (name, dims, tuple(t.shape)))


def _mat(t, name):
  """2-D float32 CUDA matrix with unit inner stride -> leading dimension."""
  if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] == 1)):
    raise ValueError("%s must be a CUDA float32 matrix with unit inner stride, got %s/%s" %
                     (name, tuple(t.shape), t.stride()))
  return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def _n(k=1):
  _ops.launch_count += k


# ---- convolutions -----------------------------------------------------------------------
def _conv1_path(Cx, Cv, Cout):
  """First-layer tensor-core path (staging copy + padded weights, exp_conv1_*): used by the AUTO and
  TMA backends for inputs the generic TMA path cannot take (state channels / Cin not a multiple of 32)."""
  if _backend != BACKEND_AUTO:
    return False
  if Cv == 0 and Cx % 32 == 0:
    return False
  return bool(_cabi.lib().exp_conv1_supported(Cx + Cv, Cout))


def _enrich32_path(Cx, Cv, Cout):
  """16 < Cin <= 32 (value network, 17 channels): enrich to 32 channels, pad the weights, then the
  generic TMA path (exp_conv_enrich32 / exp_conv_pad_weights32)."""
  if _backend != BACKEND_AUTO:
    return False
  Cin = Cx + Cv
  return (Cv > 0 or Cx % 32 != 0) and 16 < Cin <= 32 and Cout % 32 == 0


def _enrich32(x, vec, shift):
  B, IH, IW, Cx = x.shape
  Cv = 0 if vec is None else vec.shape[1]
  xs = _scratch(x.device, "enrich32", B * IH * IW * 32)[:B * IH * IW * 32].view(B, IH, IW, 32)
  _cabi.check(_cabi.lib().exp_conv_enrich32(x.data_ptr(), Cx, _p(vec), Cv, float(shift), xs.data_ptr(), B, IH, IW, _stream()),
              "exp_conv_enrich32")
  _n()
  return xs


def _conv1_stage(x, vec, shift, own=False):
  B, IH, IW, Cx = x.shape
  Cv = 0 if vec is None else vec.shape[1]
  l = _cabi.lib()
  n = l.exp_conv1_padded_input_elems(B, IH, IW)
  xp = torch.empty(n, device=x.device, dtype=torch.float32) if own else _scratch(x.device, "conv1_xp", n)
  _cabi.check(l.exp_conv1_pad_input(x.data_ptr(), Cx, _p(vec), Cv, float(shift), xp.data_ptr(), B, IH, IW, _stream()),
              "exp_conv1_pad_input")
  _n()
  return xp


def _first_path(Cx, Cv, Cout):
  """The split first layer (csrc/conv_first.cu): image channels by an exact-fp32 direct kernel, the per-image constant
  channels through a border-class table.  EXPOSURE_FIRST_LAYER=staged keeps the round-1 TMA staging path (A/B switch)."""
  return _FIRST_SPLIT and bool(_cabi.lib().exp_conv_first_supported(Cx, Cv, Cout))


_FIRST_SPLIT = os.environ.get("EXPOSURE_FIRST_LAYER", "split") != "staged"


def first_layer_split(cin, Cout=32):
  return _first_path(3, cin - 3, Cout)


def stage_first_layer(x, vec, shift, Cout):
  """The first-layer staging copy of concat(x, tile(vec)) - shift as a tensor of its own ([B, IH+2, IW+2, 16] for
  Cin <= 16, [B, IH, IW, 32] for 16 < Cin <= 32), or None when the layer needs none.  Hand it to conv_fwd /
  conv_wgrad as `staged=` (batch slices allowed) so that forward and weight gradient share ONE staging launch."""
  B, IH, IW, Cx = x.shape
  Cv = 0 if vec is None else vec.shape[1]
  if _first_path(Cx, Cv, Cout):
    return None
  if _conv1_path(Cx, Cv, Cout):
    return _conv1_stage(x, vec, shift, own=True).view(B, IH + 2, IW + 2, 16)
  if _enrich32_path(Cx, Cv, Cout):
    xs = torch.empty(B, IH, IW, 32, device=x.device, dtype=torch.float32)
    _cabi.check(_cabi.lib().exp_conv_enrich32(x.data_ptr(), Cx, _p(vec), Cv, float(shift), xs.data_ptr(), B, IH, IW, _stream()),
                "exp_conv_enrich32")
    _n()
    return xs
  return None


# Padded first-layer weights are a pure function of the weights: inside ONE optimizer step (between two Adam
# updates) they are built once per weight tensor.  Only the step schedules know that window: they run inside
# `with weight_cache():`; everywhere else every call pads afresh (a pointer is not an identity).
_wpad_cache = None


@contextlib.contextmanager
def weight_cache():
  global _wpad_cache
  outer = _wpad_cache
  _wpad_cache = {}
  try:
    yield
  finally:
    _wpad_cache = outer


def _padded_weights(W, Cin, Cout, kind):
  key = (W.data_ptr(), kind)
  ent = _wpad_cache.get(key) if _wpad_cache is not None else None
  if ent is not None:
    return ent
  l = _cabi.lib()
  cp = 16 if kind == "conv1" else 32
  Wp = torch.empty(16 * cp * Cout, device=W.device, dtype=torch.float32)
  fn = l.exp_conv1_pad_weights if kind == "conv1" else l.exp_conv_pad_weights32
  _cabi.check(fn(W.data_ptr(), Cin, Cout, Wp.data_ptr(), _stream()), "exp_conv_pad_weights")
  _n()
  if _wpad_cache is not None:
    _wpad_cache[key] = Wp
  return Wp


def conv_fwd(x, W, bias=None, vec=None, shift=0.0, mask_ref=None, post_mul=None, out=None, out2=None, staged=None):
  """4x4 stride-2 SAME conv over concat(x, tile(vec)) - shift.  Forward (bias + lrelu) or,
  with mask_ref, the forward-mode tangent (no bias, times lrelu'(mask_ref)).
  Returns y, or (y, y * post_mul) when post_mul is given."""
  _chk(x, "x", 4); _chk(W, "W", 4)
  B, IH, IW, Cx = x.shape
  Cv = 0 if vec is None else vec.shape[1]
  assert W.shape[0] == 4 and W.shape[1] == 4 and W.shape[2] == Cx + Cv, (W.shape, Cx, Cv)
  if vec is not None:
    _chk(vec, "vec", 2)
  Cout = W.shape[3]
  y = torch.empty(B, IH // 2, IW // 2, Cout, device=x.device, dtype=torch.float32) if out is None else out
  y2 = None
  if post_mul is not None:
    y2 = torch.empty_like(y) if out2 is None else out2
  mode = 0 if mask_ref is None else 1
  if staged is None and _first_path(Cx, Cv, Cout):
    with _ops._Timed("conv_fwd", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * Cx):     # the flops it executes (K = 48)
      _cabi.check(_cabi.lib().exp_conv_first_fwd(x.data_ptr(), _p(vec), Cv, float(shift), W.data_ptr(), _p(bias), _p(mask_ref),
                                                 _p(post_mul), y.data_ptr(), _p(y2), B, IH, IW, mode, _stream()),
                  "exp_conv_first_fwd")
    _n()
    return y if post_mul is None else (y, y2)
  if _conv1_path(Cx, Cv, Cout):
    l = _cabi.lib()
    with _ops._Timed("conv_fwd", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * (Cx + Cv)):
      xp = staged if staged is not None else _conv1_stage(x, vec, shift)
      Wp = _padded_weights(W, Cx + Cv, Cout, "conv1")
      _cabi.check(l.exp_conv1_fwd(xp.data_ptr(), Wp.data_ptr(), _p(bias), _p(mask_ref), _p(post_mul), y.data_ptr(), _p(y2),
                                  B, IH, IW, Cout, mode, _stream()), "exp_conv1_fwd")
    _n()
    return y if post_mul is None else (y, y2)
  if _enrich32_path(Cx, Cv, Cout):
    l = _cabi.lib()
    with _ops._Timed("conv_fwd", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * (Cx + Cv)):
      xs = staged if staged is not None else _enrich32(x, vec, shift)
      Wp = _padded_weights(W, Cx + Cv, Cout, "enrich32")
      _cabi.check(l.exp_conv_fwd(xs.data_ptr(), 32, None, 0, 0.0, Wp.data_ptr(), _p(bias), _p(mask_ref), _p(post_mul),
                                 y.data_ptr(), _p(y2), B, IH, IW, Cout, mode, _stream()), "exp_conv_fwd")
    _n()
    return y if post_mul is None else (y, y2)
  with _ops._Timed("conv_fwd", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * (Cx + Cv)):
    _cabi.check(_cabi.lib().exp_conv_fwd(x.data_ptr(), Cx, _p(vec), Cv, float(shift), W.data_ptr(), _p(bias),
                                         _p(mask_ref), _p(post_mul), y.data_ptr(), _p(y2), B, IH, IW, Cout, mode,
                                         _stream()), "exp_conv_fwd")
  _n()
  return y if post_mul is None else (y, y2)


def conv_dgrad(dy, W, in_shape, a_in=None, out=None):
  """dx[B,IH,IW,Cin] = conv^T(dy) (* lrelu'(a_in))."""
  _chk(dy, "dy", 4); _chk(W, "W", 4)
  B, IH, IW, Cin = in_shape
  Cout = W.shape[3]
  assert W.shape[2] == Cin and tuple(dy.shape) == (B, IH // 2, IW // 2, Cout)
  dx = torch.empty(B, IH, IW, Cin, device=dy.device, dtype=torch.float32) if out is None else out
  with _ops._Timed("conv_dgrad", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * Cin):
    _cabi.check(_cabi.lib().exp_conv_dgrad(dy.data_ptr(), W.data_ptr(), _p(a_in), dx.data_ptr(), B, IH, IW, Cin, Cout,
                                           _stream()), "exp_conv_dgrad")
  _n()
  return dx


def conv_first_dgrad(dy, W, Cv, in_hw, need_image=True, need_vec=True):
  """Backward of the first layer w.r.t. its input, as its producers need it: (dx_img [B,IH,IW,3], gvec [B,Cv]) -- the
  gradient of the image channels and, for every per-image constant channel, its gradient summed over the pixels."""
  _chk(dy, "dy", 4); _chk(W, "W", 4)
  B, OH, OW, Cout = dy.shape
  IH, IW = in_hw
  assert (OH, OW) == (IH // 2, IW // 2) and W.shape[2] == 3 + Cv and W.shape[3] == Cout
  dx = torch.empty(B, IH, IW, 3, device=dy.device, dtype=torch.float32) if need_image else None
  gv = torch.empty(B, Cv, device=dy.device, dtype=torch.float32) if need_vec and Cv else None
  l = _cabi.lib()
  with _ops._Timed("conv_dgrad", "gemm", 2 * B * OH * OW * Cout * 16 * 3):
    if dx is not None and gv is not None and FORK_ENABLED:
      # two independent kernels (image channels / per-image sums of the constant channels): side by side.  Both outputs
      # were allocated on the current stream above; the explicit wait (not join(): that may be deferred) orders the
      # consumer after the side stream.
      parent = torch.cuda.current_stream()
      key = (parent.device, "first_dgrad")
      side = _side_streams.get(key)
      if side is None:
        side = _side_streams[key] = torch.cuda.Stream(device=parent.device)
      side.wait_stream(parent)
      with torch.cuda.stream(side):
        _cabi.check(l.exp_conv_first_dgrad(dy.data_ptr(), W.data_ptr(), Cv, None, gv.data_ptr(), B, IH, IW, _stream()),
                    "exp_conv_first_dgrad")
      _cabi.check(l.exp_conv_first_dgrad(dy.data_ptr(), W.data_ptr(), Cv, dx.data_ptr(), None, B, IH, IW, _stream()),
                  "exp_conv_first_dgrad")
      parent.wait_stream(side)
    else:
      _cabi.check(l.exp_conv_first_dgrad(dy.data_ptr(), W.data_ptr(), Cv, _p(dx), _p(gv), B, IH, IW, _stream()),
                  "exp_conv_first_dgrad")
  _n((1 if need_image else 0) + (1 if gv is not None else 0))
  return dx, gv


def conv_wgrad(x, dy, vec=None, shift=0.0, out=None, accumulate=False, staged=None):
  """gW[4,4,Cin,Cout] (+= when accumulate) for the conv whose input was concat(x, tile(vec)) - shift."""
  _chk(x, "x", 4); _chk(dy, "dy", 4)
  B, IH, IW, Cx = x.shape
  Cv = 0 if vec is None else vec.shape[1]
  Cout = dy.shape[3]
  gW = torch.empty(4, 4, Cx + Cv, Cout, device=x.device, dtype=torch.float32) if out is None else out
  assert not accumulate or out is not None
  l = _cabi.lib()
  if staged is None and _first_path(Cx, Cv, Cout):
    ws = _workspace(x.device, l.exp_conv_first_wgrad_workspace_bytes(B, IH, IW))
    with _ops._Timed("conv_wgrad", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * Cx):
      _cabi.check(l.exp_conv_first_wgrad(x.data_ptr(), _p(vec), Cv, float(shift), dy.data_ptr(), gW.data_ptr(), B, IH, IW,
                                         int(accumulate), ws.data_ptr(), ws.numel(), _stream()), "exp_conv_first_wgrad")
    _n(2)
    return gW
  if _conv1_path(Cx, Cv, Cout):
    ws = _workspace(x.device, l.exp_conv1_wgrad_workspace_bytes(B, IH, IW, Cout))
    with _ops._Timed("conv_wgrad", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * (Cx + Cv)):
      xp = staged if staged is not None else _conv1_stage(x, vec, shift)
      _cabi.check(l.exp_conv1_wgrad(xp.data_ptr(), dy.data_ptr(), gW.data_ptr(), Cx + Cv, B, IH, IW, Cout, int(accumulate),
                                    ws.data_ptr(), ws.numel(), _stream()), "exp_conv1_wgrad")
    _n(2)
    return gW
  if _enrich32_path(Cx, Cv, Cout):
    Cin = Cx + Cv
    ws = _workspace(x.device, l.exp_conv_wgrad_workspace_bytes(B, IH, IW, 32, Cout))
    gWp = _scratch(x.device, "gwp32", 16 * 32 * Cout)[:16 * 32 * Cout].view(4, 4, 32, Cout)
    with _ops._Timed("conv_wgrad", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * Cin):
      xs = staged if staged is not None else _enrich32(x, vec, shift)
      _cabi.check(l.exp_conv_wgrad(xs.data_ptr(), 32, None, 0, 0.0, dy.data_ptr(), gWp.data_ptr(), B, IH, IW, Cout, 0,
                                   ws.data_ptr(), ws.numel(), _stream()), "exp_conv_wgrad")
      if accumulate:
        gW.add_(gWp[:, :, :Cin, :])
      else:
        gW.copy_(gWp[:, :, :Cin, :])
    _n(3)
    return gW
  nbytes = l.exp_conv_wgrad_workspace_bytes(B, IH, IW, Cx + Cv, Cout)
  ws = _workspace(x.device, nbytes)
  with _ops._Timed("conv_wgrad", "gemm", 2 * B * (IH // 2) * (IW // 2) * Cout * 16 * (Cx + Cv)):
    _cabi.check(l.exp_conv_wgrad(x.data_ptr(), Cx, _p(vec), Cv, float(shift), dy.data_ptr(), gW.data_ptr(), B, IH, IW,
                                 Cout, int(accumulate), ws.data_ptr(), ws.numel(), _stream()), "exp_conv_wgrad")
  _n(2)
  return gW


# ---- fully connected --------------------------------------------------------------------
FC_LRELU, FC_TANGENT, FC_LINEAR, FC_NOBIAS = 0, 1, 2, 3


def fc_fwd(x, W, bias=None, mode=FC_LRELU, mask_ref=None, out=None):
  ldx = _mat(x, "x"); _chk(W, "W", 2)
  M, K = x.shape
  N = W.shape[1]
  assert W.shape[0] == K
  y = torch.empty(M, N, device=x.device, dtype=torch.float32) if out is None else out
  ldy = _mat(y, "y")
  ldm = _mat(mask_ref, "mask_ref") if mask_ref is not None else 0
  l = _cabi.lib()
  ws = _workspace(x.device, l.exp_fc_workspace_bytes(M, K, N))
  _cabi.check(l.exp_fc_fwd(x.data_ptr(), ldx, W.data_ptr(), _p(bias), _p(mask_ref), ldm, y.data_ptr(), ldy, M, K, N,
                           mode, ws.data_ptr(), ws.numel(), _stream()), "exp_fc_fwd")
  _n(2)
  return y


def fc_dgrad(dy, W, mul_act=None, mul_plain=None, out=None, accumulate=False):
  """dx = (out if accumulate else 0) + dy W^T * lrelu'(mul_act) * mul_plain."""
  ldy = _mat(dy, "dy"); _chk(W, "W", 2)
  M, N = dy.shape
  K = W.shape[0]
  assert W.shape[1] == N
  dx = torch.empty(M, K, device=dy.device, dtype=torch.float32) if out is None else out
  lddx = _mat(dx, "dx")
  ldmul = 0
  for m in (mul_act, mul_plain):
    if m is not None:
      l_ = _mat(m, "mul")
      assert ldmul in (0, l_), "both multipliers must share a leading dimension"
      ldmul = l_
  _cabi.check(_cabi.lib().exp_fc_dgrad(dy.data_ptr(), ldy, W.data_ptr(), _p(mul_act), _p(mul_plain), ldmul,
                                       dx.data_ptr(), lddx, M, K, N, int(accumulate), _stream()), "exp_fc_dgrad")
  _n()
  return dx


def fc_wgrad(x, dy, out=None, accumulate=False):
  ldx = _mat(x, "x"); ldy = _mat(dy, "dy")
  M, K = x.shape
  N = dy.shape[1]
  gW = torch.empty(K, N, device=x.device, dtype=torch.float32) if out is None else out
  _chk(gW, "gW", 2)
  _cabi.check(_cabi.lib().exp_fc_wgrad(x.data_ptr(), ldx, dy.data_ptr(), ldy, gW.data_ptr(), M, K, N, int(accumulate),
                                       _stream()), "exp_fc_wgrad")
  _n()
  return gW


def colsum(a, batch=1, out=None):
  """out[batch, cols] = column sums of a viewed as [batch, rows, cols] (cols = last dim)."""
  _chk(a, "a")
  cols = a.shape[-1]
  rows = a.numel() // cols // batch
  o = torch.empty((batch, cols) if batch > 1 else (cols,), device=a.device, dtype=torch.float32) if out is None else out
  l = _cabi.lib()
  nbytes = l.exp_colsum_workspace_bytes(batch, rows, cols)
  ws = _workspace(a.device, nbytes) if nbytes else None
  _cabi.check(l.exp_colsum(a.data_ptr(), batch, rows, cols, o.data_ptr(), _p(ws), ws.numel() if ws is not None else 0,
                           _stream()), "exp_colsum")
  _n(2 if nbytes else 1)
  return o


# ---- per-image heads --------------------------------------------------------------------
def stats_fwd(img):
  _chk(img, "img", 4)
  B, H, W, _ = img.shape
  s = torch.empty(B, 3, device=img.device, dtype=torch.float32)
  _cabi.check(_cabi.lib().exp_stats_fwd(img.data_ptr(), s.data_ptr(), B, H, W, _stream()), "exp_stats_fwd")
  _n()
  return s


def stats_bwd(img, stats, g_stat, g_direct=None, out=None):
  _chk(img, "img", 4); _chk(g_stat, "g_stat", 2)
  B, H, W, _ = img.shape
  g = torch.empty_like(img) if out is None else out
  _cabi.check(_cabi.lib().exp_stats_bwd(img.data_ptr(), stats.data_ptr(), g_stat.data_ptr(), _p(g_direct), g.data_ptr(),
                                        B, H, W, _stream()), "exp_stats_bwd")
  _n()
  return g


def stats_jvp(img, stats, u):
  _chk(img, "img", 4); _chk(u, "u", 4)
  B, H, W, _ = img.shape
  d = torch.empty(B, 3, device=img.device, dtype=torch.float32)
  _cabi.check(_cabi.lib().exp_stats_jvp(img.data_ptr(), stats.data_ptr(), u.data_ptr(), d.data_ptr(), B, H, W, _stream()),
              "exp_stats_jvp")
  _n()
  return d


def _progress_dev(progress, dev):
  """progress may be a python float (uploaded) or a device scalar tensor (graph-friendly)."""
  if torch.is_tensor(progress):
    return progress
  return torch.full((1,), float(progress), device=dev)


def policy_head_fwd(logits, noise, states, is_train, progress, cfg):
  _chk(logits, "logits", 2); _chk(noise, "noise"); _chk(states, "states", 2)
  progress = _progress_dev(progress, logits.device)
  B, n = logits.shape
  dev = logits.device
  pdf = torch.empty(B, n, device=dev)
  ids = torch.empty(B, dtype=torch.int32, device=dev)
  sur = torch.empty(B, device=dev); ent = torch.empty(B, device=dev); pen = torch.empty(B, device=dev)
  ns = torch.empty_like(states)
  _cabi.check(_cabi.lib().exp_policy_head_fwd(
      logits.data_ptr(), noise.data_ptr(), states.data_ptr(), B, n, states.shape[1], int(is_train), int(cfg.test_steps),
      float(cfg.exploration), float(cfg.exploration_penalty), float(cfg.filter_usage_penalty), progress.data_ptr(),
      pdf.data_ptr(), ids.data_ptr(), sur.data_ptr(), ent.data_ptr(), pen.data_ptr(), ns.data_ptr(), _stream()),
      "exp_policy_head_fwd")
  _n()
  return pdf, ids, sur, ent, pen, ns


def policy_head_bwd(logits, ids, g_surrogate, g_penalty, progress, cfg):
  B, n = logits.shape
  g = torch.empty_like(logits)
  progress = _progress_dev(progress, logits.device)
  _cabi.check(_cabi.lib().exp_policy_head_bwd(logits.data_ptr(), ids.data_ptr(), g_surrogate.data_ptr(),
                                              g_penalty.data_ptr(), B, n, float(cfg.exploration),
                                              float(cfg.exploration_penalty), progress.data_ptr(), g.data_ptr(), _stream()),
              "exp_policy_head_bwd")
  _n()
  return g


def overexposure_fwd(img):
  _chk(img, "img", 4)
  B, H, W, _ = img.shape
  pen = torch.empty(B, device=img.device)
  _cabi.check(_cabi.lib().exp_overexposure_fwd(img.data_ptr(), pen.data_ptr(), B, H, W, _stream()), "exp_overexposure_fwd")
  _n()
  return pen


def overexposure_bwd(img, g_pen, g_in=None, out=None):
  B, H, W, _ = img.shape
  g = torch.empty_like(img) if out is None else out
  _cabi.check(_cabi.lib().exp_overexposure_bwd(img.data_ptr(), g_pen.data_ptr(), _p(g_in), g.data_ptr(), B, H, W, _stream()),
              "exp_overexposure_bwd")
  _n()
  return g


def rl_losses(fake_logit, fake_input_logit, old_value, new_value, penalty, surrogate, new_states, cfg):
  """Returns (seeds[5,B], losses[2]); see exp_rl_losses."""
  B = fake_logit.numel()
  dev = fake_logit.device
  seeds = torch.empty(5, B, device=dev)
  losses = torch.empty(2, device=dev)
  _cabi.check(_cabi.lib().exp_rl_losses(
      fake_logit.data_ptr(), fake_input_logit.data_ptr(), old_value.data_ptr(), new_value.data_ptr(), penalty.data_ptr(),
      surrogate.data_ptr(), new_states.data_ptr(), B, new_states.shape[1], float(cfg.all_reward),
      float(cfg.critic_logit_multiplier), float(cfg.discount_factor), float(cfg.parameter_lr_mul),
      int(cfg.maximum_trajectory_length), int(bool(cfg.use_penalty)), seeds.data_ptr(), losses.data_ptr(), _stream()),
      "exp_rl_losses")
  _n()
  return seeds, losses


def interpolate(real, fake, alpha, out=None):
  B = real.shape[0]
  n = real.numel() // B
  o = torch.empty_like(real) if out is None else out
  _cabi.check(_cabi.lib().exp_interpolate(real.data_ptr(), fake.data_ptr(), alpha.data_ptr(), o.data_ptr(), B, n, _stream()),
              "exp_interpolate")
  _n()
  return o


def gp_scale(g, lam, batch_for_mean=None, out=None):
  """Returns (u, norm): u = g * lam * 2 max(norm-1,0) / (B norm)."""
  B = g.shape[0]
  n = g.numel() // B
  u = torch.empty_like(g) if out is None else out
  norm = torch.empty(B, device=g.device)
  _cabi.check(_cabi.lib().exp_gp_scale(g.data_ptr(), u.data_ptr(), norm.data_ptr(), float(lam),
                                       int(batch_for_mean or B), n, _stream()), "exp_gp_scale")
  _n()
  return u, norm


def adam(params, grads, m, v, hyper, beta1, beta2, eps=1e-8, grad_scale=1.0):
  """In-place fused Adam on flat buffers; hyper[0] = lr_t (device scalar)."""
  for t in (params, grads, m, v):
    _chk(t, "adam buffer")
  _cabi.check(_cabi.lib().exp_adam(params.data_ptr(), grads.data_ptr(), m.data_ptr(), v.data_ptr(), hyper.data_ptr(),
                                   float(beta1), float(beta2), float(eps), float(grad_scale), params.numel(), _stream()),
              "exp_adam")
  _n()


# ---- fused bookkeeping kernels of the train step (csrc/train_glue.cu) -----------------------------------------
import ctypes as _ct


def _iarr(vals):
  return (_ct.c_int * len(vals))(*[int(v) for v in vals])


def critic_inputs(real, fake, alpha, out=None):
  """X [3B,...] = real | fake | real + alpha (fake - real) (net.py:174-179 + the batch concat), one launch."""
  _chk(real, "real"); _chk(fake, "fake")
  B = real.shape[0]
  n = real.numel() // B
  X = torch.empty((3 * B,) + tuple(real.shape[1:]), device=real.device, dtype=torch.float32) if out is None else out
  _cabi.check(_cabi.lib().exp_critic_inputs(real.data_ptr(), fake.data_ptr(), alpha.data_ptr(), X.data_ptr(), B, n, _stream()),
              "exp_critic_inputs")
  _n()
  return X


def critic_scalars(logits, norm, lam, ema_state=None, decay=0.99):
  """out[5] = emd, gradient penalty, critic_gradient_norm, c_loss, c_average; advances the zero-debiased moving
  average in ema_state [3] when given (net.py:164-168, 185-187, 268-269)."""
  B = norm.numel()
  out = torch.empty(8, device=logits.device, dtype=torch.float32)
  _cabi.check(_cabi.lib().exp_critic_scalars(logits.data_ptr(), norm.data_ptr(), B, float(lam), _p(ema_state), float(decay),
                                             out.data_ptr(), _stream()), "exp_critic_scalars")
  _n()
  return out


class HeadsLayout:
  """Where the n_heads fc2 layers of the filter heads live inside the generator's flat parameter buffer."""

  def __init__(self, store, names, dims, npar, fc1, nmask):
    self.store, self.n = store, len(names)
    self.w_off = _iarr([store.offsets[n + "/weights"][0] for n in names])
    self.b_off = _iarr([store.offsets[n + "/biases"][0] for n in names])
    self.dims, self.npar = _iarr(dims), _iarr(npar)
    self.fc1, self.nmask, self.ostride = fc1, nmask, max(dims)


def heads_fc2_fwd(L, H):
  B = H.shape[0]
  ldh = _mat(H, "H")
  O = torch.empty(B, L.n, L.ostride, device=H.device, dtype=torch.float32)
  _cabi.check(_cabi.lib().exp_heads_fc2_fwd(L.store.flat.data_ptr(), L.w_off, L.b_off, L.dims, L.npar, L.n, L.fc1, L.nmask,
                                            H.data_ptr(), ldh, O.data_ptr(), L.ostride, B, _stream()), "exp_heads_fc2_fwd")
  _n()
  return O


def heads_select(L, O, ids, selstride, want_mask):
  B = O.shape[0]
  sel = torch.empty(B, selstride, device=O.device, dtype=torch.float32)
  msel = torch.empty(B, L.nmask, device=O.device, dtype=torch.float32) if want_mask else None
  _cabi.check(_cabi.lib().exp_heads_select(O.data_ptr(), L.ostride, ids.data_ptr(), L.npar, L.n, L.nmask, sel.data_ptr(), selstride,
                                           _p(msel), B, _stream()), "exp_heads_select")
  _n()
  return sel, msel


def heads_fc2_bwd(L, H, ids, gsel, gmsel=None):
  """Returns dH [B, n_heads * fc1]; overwrites the fc2 weight / bias gradients of every head in the store."""
  B = H.shape[0]
  ldh = _mat(H, "H")
  _chk(gsel, "gsel", 2)
  dH = torch.empty_like(H)
  assert dH.stride(0) == ldh
  _cabi.check(_cabi.lib().exp_heads_fc2_bwd(L.store.flat.data_ptr(), L.store.grad.data_ptr(), L.w_off, L.b_off, L.dims, L.npar, L.n,
                                            L.fc1, L.nmask, H.data_ptr(), ldh, ids.data_ptr(), gsel.data_ptr(), gsel.shape[1],
                                            _p(gmsel), dH.data_ptr(), B, _stream()), "exp_heads_fc2_bwd")
  _n()
  return dH


_zero_ws = {}


def _zero_workspace(dev, nbytes):
  """Zero-initialised, self-cleaning workspace per (device, stream): ticket counters of exp_colsum_multi."""
  key = (dev, torch.cuda.current_stream().cuda_stream)
  ws = _zero_ws.get(key)
  if ws is None or ws.numel() < nbytes:
    ws = torch.zeros(max(int(nbytes), 1 << 16), dtype=torch.uint8, device=dev)
    _zero_ws[key] = ws
  return ws


def colsum_multi(tasks):
  """tasks: list of (src [rows, cols] (any leading dims), dst [cols], accumulate) -- up to 8 bias-gradient column sums per
  launch (deterministic)."""
  l = _cabi.lib()
  for i in range(0, len(tasks), 8):
    part = tasks[i:i + 8]
    n = len(part)
    rows = _iarr([t[0].numel() // t[0].shape[-1] for t in part])
    cols = _iarr([t[0].shape[-1] for t in part])
    acc = _iarr([1 if t[2] else 0 for t in part])
    for t in part:
      _chk(t[0], "colsum source")
      assert t[1].is_contiguous() and t[1].numel() == t[0].shape[-1]
    src = (_ct.c_void_p * n)(*[t[0].data_ptr() for t in part])
    dst = (_ct.c_void_p * n)(*[t[1].data_ptr() for t in part])
    ws = _zero_workspace(part[0][0].device, l.exp_colsum_multi_workspace_bytes(rows, cols, n))
    _cabi.check(l.exp_colsum_multi(src, dst, rows, cols, acc, n, ws.data_ptr(), ws.numel(), _stream()), "exp_colsum_multi")
    _n()


def set_floats(dst, values):
  """dst[:len(values)] = values (<= 16 floats) in ONE launch, values passed by value (exp_set_floats)."""
  vals = (_ct.c_float * len(values))(*[float(v) for v in values])
  _cabi.check(_cabi.lib().exp_set_floats(dst.data_ptr(), _ct.cast(vals, _ct.c_void_p), len(values), _stream()), "exp_set_floats")
  _n()


def stats_bwd_gin(img, stats, g_in, out=None):
  """dL/dimages from the layer-1 input gradient g_in [n,H,W,cin] (image channels + J_stats^T of the statistic channels)."""
  _chk(img, "img", 4); _chk(g_in, "g_in", 4)
  B, H, W, _ = img.shape
  g = torch.empty_like(img) if out is None else out
  _cabi.check(_cabi.lib().exp_stats_bwd_gin(img.data_ptr(), stats.data_ptr(), g_in.data_ptr(), g_in.shape[3], g.data_ptr(), B, H, W,
                                            _stream()), "exp_stats_bwd_gin")
  _n()
  return g


# ---- device-side replay memory + per-step random draws (csrc/replay.cu) -------------------------------------------
def replay_draw_generator(pool_states, B, seed, ctl, batch_src, rest_src):
  P, S = pool_states.shape
  _cabi.check(_cabi.lib().exp_replay_draw_generator(pool_states.data_ptr(), S, P, B, int(seed), ctl.data_ptr(), batch_src.data_ptr(),
                                                    rest_src.data_ptr(), _stream()), "exp_replay_draw_generator")
  _n()


def replay_replace(new_states, P, max_traj_len, keep_prob, seed, ctl, rest_src, new_pool_src):
  B, S = new_states.shape
  _cabi.check(_cabi.lib().exp_replay_replace(new_states.data_ptr(), S, P, B, int(max_traj_len), float(keep_prob), int(seed),
                                             ctl.data_ptr(), rest_src.data_ptr(), new_pool_src.data_ptr(), _stream()),
              "exp_replay_replace")
  _n()


def replay_draw_critic(pool_states, B, seed, ctl, batch_src):
  P, S = pool_states.shape
  _cabi.check(_cabi.lib().exp_replay_draw_critic(pool_states.data_ptr(), S, P, B, int(seed), ctl.data_ptr(), batch_src.data_ptr(),
                                                 _stream()), "exp_replay_draw_critic")
  _n()


def gather_rows(src, idx, out):
  """out[i] = src[idx[i]] over the leading dimension (idx: CUDA int64 [n])."""
  n = idx.numel()
  row = src.numel() // src.shape[0]
  assert src.is_contiguous() and out.is_contiguous() and out.numel() == n * row and idx.dtype == torch.int64
  _cabi.check(_cabi.lib().exp_gather_rows(src.data_ptr(), idx.data_ptr(), out.data_ptr(), n, row, _stream()), "exp_gather_rows")
  _n()
  return out


def train_draws(seed, ctl, uniform=None, mask=None, keep=0.5):
  """uniform[...] ~ U[0,1), mask[...] = floor(keep + U) / keep (tf.nn.dropout multiplier), one launch (+ the counter bump)."""
  nu = uniform.numel() if uniform is not None else 0
  nm = mask.numel() if mask is not None else 0
  _cabi.check(_cabi.lib().exp_train_draws(int(seed), ctl.data_ptr(), _p(uniform), nu, _p(mask), nm, float(keep), _stream()),
              "exp_train_draws")
  _n(2)


# 0 = TMA-fed tcgen05 engine wherever the shape allows it (the product path); 1 = exact-fp32 CUDA-core engine
# everywhere (the tests' A/B switch).  BACKEND_TCGEN05_TMA is the old name of 0.
BACKEND_AUTO, BACKEND_CUDA_CORES = 0, 1
BACKEND_TCGEN05_TMA = BACKEND_AUTO
_backend = int(os.environ.get("EXPOSURE_GEMM_BACKEND") or 0)     # mirrors _cabi.lib()'s initial setting


def set_gemm_backend(backend):
  """Process-wide GEMM backend of the conv / FC primitives (exp_set_gemm_backend)."""
  global _backend
  _cabi.check(_cabi.lib().exp_set_gemm_backend(int(backend)), "exp_set_gemm_backend")
  _backend = int(backend)
