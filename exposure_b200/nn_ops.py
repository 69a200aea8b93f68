"""torch-tensor front end of the network primitives of the C ABI (conv / FC / colsum).

Plumbing only: device pointers, current stream, output allocation.  See
include/exposure_b200.h for the semantics of every call."""
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


def _p(t):
  return None if t is None else t.data_ptr()


def _chk(t, name, dims=None):
  if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
    raise ValueError("%s must be a contiguous CUDA float32 tensor" % name)
  if dims is not None and t.dim() != dims:
    raise ValueError("%s must have %d dims, got %s" % (name, dims, tuple(t.shape)))


def conv_fwd(x, W, bias=None, vec=None, shift=0.0, mask_ref=None, post_mul=None, out=None, out2=None):
  """4x4 stride-2 SAME conv over concat(x, tile(vec)) - shift.  Forward (bias + lrelu) or,
  with mask_ref, the forward-mode tangent (no bias, times lrelu'(mask_ref)).
  Returns y, or (y, y * post_mul) when post_mul is given."""
  _chk(x, "x", 4); _chk(W, "W", 4)
  B, IH, IW, Cx = x.shape
  Cv = 0 if vec is None else vec.shape[1]
  assert W.shape[0] == 4 and W.shape[1] == 4 and W.shape[2] == Cx + Cv, (W.shape, Cx, Cv)
  Cout = W.shape[3]
  y = torch.empty(B, IH // 2, IW // 2, Cout, device=x.device, dtype=torch.float32) if out is None else out
  y2 = None
  if post_mul is not None:
    y2 = torch.empty_like(y) if out2 is None else out2
  mode = 0 if mask_ref is None else 1
  _cabi.check(_cabi.lib().exp_conv_fwd(x.data_ptr(), Cx, _p(vec), Cv, float(shift), W.data_ptr(), _p(bias),
                                       _p(mask_ref), _p(post_mul), y.data_ptr(), _p(y2), B, IH, IW, Cout, mode,
                                       _stream()), "exp_conv_fwd")
  _ops.launch_count += 1
  return y if post_mul is None else (y, y2)


def conv_dgrad(dy, W, in_shape, a_in=None, out=None):
  """dx[B,IH,IW,Cin] = conv^T(dy) (* lrelu'(a_in))."""
  _chk(dy, "dy", 4); _chk(W, "W", 4)
  B, IH, IW, Cin = in_shape
  Cout = W.shape[3]
  assert W.shape[2] == Cin and dy.shape == (B, IH // 2, IW // 2, Cout)
  dx = torch.empty(B, IH, IW, Cin, device=dy.device, dtype=torch.float32) if out is None else out
  _cabi.check(_cabi.lib().exp_conv_dgrad(dy.data_ptr(), W.data_ptr(), _p(a_in), dx.data_ptr(), B, IH, IW, Cin, Cout,
                                         _stream()), "exp_conv_dgrad")
  _ops.launch_count += 1
  return dx


def conv_wgrad(x, dy, vec=None, shift=0.0, out=None):
  """gW[4,4,Cin,Cout] for the conv whose input was concat(x, tile(vec)) - shift."""
  _chk(x, "x", 4); _chk(dy, "dy", 4)
  B, IH, IW, Cx = x.shape
  Cv = 0 if vec is None else vec.shape[1]
  Cout = dy.shape[3]
  gW = torch.empty(4, 4, Cx + Cv, Cout, device=x.device, dtype=torch.float32) if out is None else out
  l = _cabi.lib()
  nbytes = l.exp_conv_wgrad_workspace_bytes(B, IH, IW, Cx + Cv, Cout)
  ws = _workspace(x.device, nbytes)
  _cabi.check(l.exp_conv_wgrad(x.data_ptr(), Cx, _p(vec), Cv, float(shift), dy.data_ptr(), gW.data_ptr(), B, IH, IW,
                               Cout, ws.data_ptr(), ws.numel(), _stream()), "exp_conv_wgrad")
  _ops.launch_count += 2
  return gW


FC_LRELU, FC_TANGENT, FC_LINEAR, FC_NOBIAS = 0, 1, 2, 3


def fc_fwd(x, W, bias=None, mode=FC_LRELU, mask_ref=None, out=None):
  _chk(x, "x", 2); _chk(W, "W", 2)
  M, K = x.shape
  N = W.shape[1]
  assert W.shape[0] == K
  y = torch.empty(M, N, device=x.device, dtype=torch.float32) if out is None else out
  l = _cabi.lib()
  ws = _workspace(x.device, l.exp_fc_workspace_bytes(M, K, N))
  _cabi.check(l.exp_fc_fwd(x.data_ptr(), W.data_ptr(), _p(bias), _p(mask_ref), y.data_ptr(), M, K, N, mode,
                           ws.data_ptr(), ws.numel(), _stream()), "exp_fc_fwd")
  _ops.launch_count += 2
  return y


def fc_dgrad(dy, W, mul=None, mul_mode=0, out=None):
  _chk(dy, "dy", 2); _chk(W, "W", 2)
  M, N = dy.shape
  K = W.shape[0]
  dx = torch.empty(M, K, device=dy.device, dtype=torch.float32) if out is None else out
  _cabi.check(_cabi.lib().exp_fc_dgrad(dy.data_ptr(), W.data_ptr(), _p(mul), mul_mode, dx.data_ptr(), M, K, N,
                                       _stream()), "exp_fc_dgrad")
  _ops.launch_count += 1
  return dx


def fc_wgrad(x, dy, out=None):
  _chk(x, "x", 2); _chk(dy, "dy", 2)
  M, K = x.shape
  N = dy.shape[1]
  gW = torch.empty(K, N, device=x.device, dtype=torch.float32) if out is None else out
  _cabi.check(_cabi.lib().exp_fc_wgrad(x.data_ptr(), dy.data_ptr(), gW.data_ptr(), M, K, N, _stream()), "exp_fc_wgrad")
  _ops.launch_count += 1
  return gW


def colsum(a, out=None):
  """Column sums of a [rows, cols] (any leading dims are flattened into rows)."""
  _chk(a, "a")
  cols = a.shape[-1]
  rows = a.numel() // cols
  o = torch.empty(cols, device=a.device, dtype=torch.float32) if out is None else out
  _cabi.check(_cabi.lib().exp_colsum(a.data_ptr(), rows, cols, o.data_ptr(), _stream()), "exp_colsum")
  _ops.launch_count += 1
  return o
