"""Minimal eager stand-in for the TensorFlow 1.x Python API -- GOLDEN-VECTOR TOOLING ONLY.

Purpose: let the reference's OWN Python (filters.py, agent.py, critics.py, pdf_sample_layer.py,
util.py under /root/reference) execute in the build container, where TensorFlow 1.6 does not exist,
so that tests/golden/make_reference_golden.py can record what the reference's code computes.  Only the
API subset those files touch is provided; every op runs eagerly on torch CPU tensors in the dtype of
its inputs, so torch autograd stands in for tf.gradients.

What this is NOT: TensorFlow.  The primitives below restate TF 1.6's documented kernel behaviour
(tie rules of maximum/minimum/clip_by_value, population variance in nn.moments, the RGB<->HSV
functors of tensorflow/core/kernels/colorspace_op.h, floor(keep+U)/keep dropout).  The op ORDER and
every formula above the primitives come from the reference source itself.

Nothing in exposure_b200/, bench.py or the GPU tests imports this package."""
import contextlib

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64

# --------------------------------------------------------------------------------------------
# tensors: plain torch tensors; TF spells the static shape get_shape()
torch.Tensor.get_shape = lambda self: self.shape


def _accept_ndarray(name):
  """`tensor <op> ndarray` (filters.py:228 multiplies by a numpy mask) must stay on the autograd tape:
  torch would otherwise hand the expression to numpy."""
  orig = getattr(torch.Tensor, name)

  def op(self, other):
    if isinstance(other, np.ndarray):
      other = torch.from_numpy(other).to(self.dtype)
    return orig(self, other)

  setattr(torch.Tensor, name, op)


# TF tensors are immutable: `a *= b` rebinds the name (filters.py:232, 142, 409).  torch would update in
# place and invalidate the tape, so the augmented assignments are made out-of-place.
torch.Tensor.__imul__ = lambda self, other: self * other
torch.Tensor.__iadd__ = lambda self, other: self + other
torch.Tensor.__isub__ = lambda self, other: self - other
torch.Tensor.__itruediv__ = lambda self, other: self / other

for _n in ("__mul__", "__rmul__", "__add__", "__radd__", "__sub__", "__rsub__", "__truediv__", "__rtruediv__"):
  _accept_ndarray(_n)

_DTYPES = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32,
           np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64, np.dtype("int32"): torch.int32}

# dtype the float "constants" are created in: the generating script switches it to run the reference
# code in fp64 (arbiter) as well as in its native fp32
_state = {"float": torch.float32}


def set_float_dtype(dt):
  _state["float"] = dt


def _dt(d):
  if d is None:
    return _state["float"]
  d = _DTYPES.get(d, d)
  if d in (torch.float32, torch.float64):
    return _state["float"]
  return d


def _t(x, like=None):
  if isinstance(x, torch.Tensor):
    return x
  if isinstance(x, np.ndarray):
    t = torch.from_numpy(x)
    return t.to(_state["float"]) if t.is_floating_point() else t
  dt = like.dtype if like is not None else _state["float"]
  return torch.tensor(x, dtype=dt)


def constant(value, dtype=None):
  return _t(np.asarray(value)) if dtype is None else _t(np.asarray(value)).to(_dt(dtype))


def zeros(shape, dtype=float32):
  return torch.zeros(tuple(shape), dtype=_dt(dtype))


def ones(shape, dtype=float32):
  return torch.ones(tuple(shape), dtype=_dt(dtype))


class _Placeholder:
  """tf.placeholder: only its static shape is ever read by the code run here (replay_memory.py:16-40,112-116)."""

  def __init__(self, dtype, shape=None, name=None):
    self.dtype, self.shape, self.name = dtype, shape, name

  def get_shape(self):
    return self.shape


# Graph-construction code (net.py GAN.__init__) runs eagerly when its placeholders already hold the values
# the session would be fed: `feeds[name]` -> the tensor returned for tf.placeholder(..., name=name).
feeds = {}


def placeholder(dtype, shape=None, name=None):
  if name in feeds:
    v = feeds[name]
    return v if isinstance(v, torch.Tensor) else torch.tensor(v, dtype=_dt(dtype))
  return _Placeholder(dtype, shape, name)


def equal(x, y):
  return x == y


class _Obj:
  """attribute bag for configuration protos / handles nobody reads back"""

  def __getattr__(self, k):
    o = _Obj()
    object.__setattr__(self, k, o)
    return o


def ConfigProto(**k):
  return _Obj()


class Session:

  def __init__(self, config=None):
    self.graph = None

  def run(self, *a, **k):
    return None


class Variable:
  """non-trainable counters (net.py:216, 229, 241); trainable variables come from get_variable"""

  def __init__(self, initial_value=0, trainable=False, dtype=None, name=None):
    self.value = initial_value


class GraphKeys:
  TRAINABLE_VARIABLES = "trainable_variables"


def get_collection(key, scope=None):
  assert key == GraphKeys.TRAINABLE_VARIABLES
  return [_NamedVar(k, v) for k, v in _store.vars.items() if scope is None or k.startswith(scope + "/") or k == scope]


class _NamedVar:
  def __init__(self, name, tensor):
    self.name, self.tensor = name, tensor


@contextlib.contextmanager
def control_dependencies(ops):
  yield


def group(*ops):
  return ops


def assign(var, value):
  return ("assign", var, value)


def global_variables_initializer():
  return None


class _Summary:
  @staticmethod
  def scalar(name, value):
    return None

  @staticmethod
  def merge_all():
    return None

  @staticmethod
  def FileWriter(*a, **k):
    return _Obj()


summary = _Summary


# --------------------------------------------------------------------------------------------
# scopes and variables (tf.variable_scope / tf.get_variable semantics that decide checkpoint names)
class _Store:

  def __init__(self):
    self.stack = []          # current scope path
    self.counts = {}         # full scope name -> times opened (default_name uniquification)
    self.vars = {}           # full variable name -> tensor
    self.created = []        # names in creation order
    self.init_seed = 0


_store = _Store()


def reset_variables(values=None, seed=0):
  """Start a fresh 'graph': clears scope counters; `values` (name -> array/tensor) pre-loads variables,
  e.g. the shipped checkpoint.  Missing variables come from `initializer[0]`."""
  _store.stack, _store.counts, _store.created = [], {}, []
  _store.vars = {}
  _store.init_seed = seed
  for k, v in (values or {}).items():
    _store.vars[k] = torch.as_tensor(np.asarray(v)).to(_state["float"]).clone().requires_grad_(True)


def variables():
  return _store.vars


def created_variables():
  return list(_store.created)


def _path(name):
  return "/".join(_store.stack + [name]) if name else "/".join(_store.stack)


class _Scope:

  def __init__(self, full):
    self.name = full

  def reuse_variables(self):
    pass                     # the store always returns an existing variable of that name


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, reuse=None, values=None):
  if name_or_scope is None:
    # tf.variable_scope(None, default_name=...): unique within the parent scope ("Conv", "Conv_1", ...)
    base = _path(default_name)
    n = _store.counts.get(base, 0)
    _store.counts[base] = n + 1
    name = default_name if n == 0 else "%s_%d" % (default_name, n)
  else:
    name = name_or_scope if isinstance(name_or_scope, str) else name_or_scope.name.split("/")[-1]
    full = _path(name)
    _store.counts[full] = _store.counts.get(full, 0) + 1
    # VariableScopeStore.close_variable_subscopes: re-entering a scope restarts its children's numbering
    for k in list(_store.counts):
      if k.startswith(full + "/"):
        del _store.counts[k]
  _store.stack.append(name)
  try:
    yield _Scope(_path(None))
  finally:
    _store.stack.pop()


@contextlib.contextmanager
def name_scope(name=None, default_name=None, values=None):
  yield name


def _default_initializer(full_name, shape):
  g = torch.Generator().manual_seed(_store.init_seed)
  return torch.randn(tuple(shape), generator=g, dtype=torch.float64) * 0.02


initializer = [_default_initializer]     # callable(full_name, shape) -> float64 tensor; the script plugs in its own


def get_variable(name, shape):
  full = _path(name)
  if full not in _store.vars:
    _store.vars[full] = initializer[0](full, shape).to(_state["float"]).requires_grad_(True)
  v = _store.vars[full]
  assert tuple(v.shape) == tuple(shape), (full, tuple(v.shape), tuple(shape))
  if full not in _store.created:
    _store.created.append(full)
  return v


# --------------------------------------------------------------------------------------------
# element-wise math.  maximum/minimum: TF's _MaximumMinimumGrad sends the gradient to x where
# x >= y (x <= y for minimum) and to y elsewhere -- exactly what where() differentiates to.
def _pair(x, y):
  x = _t(x, y if isinstance(y, torch.Tensor) else None)
  y = _t(y, x)
  if y.dim() == 0:
    y = y.to(x.dtype)
  return torch.broadcast_tensors(x, y)


def maximum(x, y):
  x, y = _pair(x, y)
  return torch.where(x >= y, x, y)


def minimum(x, y):
  x, y = _pair(x, y)
  return torch.where(x <= y, x, y)


def clip_by_value(t, clip_value_min, clip_value_max):
  # TF 1.6 clip_ops.clip_by_value: minimum(t, max) then maximum(., min)
  return maximum(minimum(t, clip_value_max), clip_value_min)


exp = torch.exp
log = torch.log
cos = torch.cos
tanh = torch.tanh
sigmoid = torch.sigmoid
sqrt = torch.sqrt
abs = torch.abs                                        # noqa: A001  (TF's name)


def pow(x, y):                                         # noqa: A001
  return torch.pow(x, y)


def less(x, y):
  return x < y


def cast(x, dtype):
  return x.to(_dt(dtype))


def to_float(x):
  return x.to(_state["float"])


def _axes(axis, reduction_indices):
  a = axis if axis is not None else reduction_indices
  if a is None:
    return None
  return tuple(a) if isinstance(a, (list, tuple)) else (a,)


def reduce_sum(x, axis=None, keep_dims=False, reduction_indices=None):
  a = _axes(axis, reduction_indices)
  return x.sum() if a is None else x.sum(dim=a, keepdim=keep_dims)


def reduce_mean(x, axis=None, keep_dims=False, reduction_indices=None):
  a = _axes(axis, reduction_indices)
  return x.mean() if a is None else x.mean(dim=a, keepdim=keep_dims)


def reduce_max(x, axis=None, keep_dims=False, reduction_indices=None):
  # _MinOrMaxGrad: gradient split evenly among the tied maxima == torch.amax
  a = _axes(axis, reduction_indices)
  return x.amax() if a is None else x.amax(dim=a, keepdim=keep_dims)


def reduce_min(x, axis=None, keep_dims=False, reduction_indices=None):
  a = _axes(axis, reduction_indices)
  return x.amin() if a is None else x.amin(dim=a, keepdim=keep_dims)


def argmax(x, axis=None):
  return torch.argmax(x, dim=axis)


def cumsum(x, axis=0, exclusive=False, reverse=False):
  assert not reverse
  c = torch.cumsum(x, dim=axis)
  if exclusive:                                        # shift right by one, zero first
    c = torch.cat([torch.zeros_like(c.narrow(axis, 0, 1)), c.narrow(axis, 0, c.shape[axis] - 1)], dim=axis)
  return c


def one_hot(indices, depth, dtype=float32):
  # out-of-range indices (e.g. -1) give an all-zero row, like TF
  r = torch.arange(depth, dtype=torch.int64)
  return (indices.to(torch.int64)[..., None] == r).to(_dt(dtype))


def reshape(x, shape):
  return x.reshape(tuple(int(s) for s in shape))


def concat(values, axis):
  return torch.cat(list(values), dim=axis)


def stack(values, axis=0):
  return torch.stack(list(values), dim=axis)


def tile(x, multiples):
  return x.repeat(*[int(m) for m in multiples])


def stop_gradient(x):
  return x.detach()


def gradients(ys, xs, grad_ys=None):
  single = not isinstance(xs, (list, tuple))
  xs_ = [xs] if single else list(xs)
  ys_ = ys if isinstance(ys, (list, tuple)) else [ys]
  total = sum(y.sum() if grad_ys is None else (y * grad_ys).sum() for y in ys_)
  g = torch.autograd.grad(total, xs_, create_graph=True, allow_unused=True)
  return list(g)


# --------------------------------------------------------------------------------------------
class _NN:
  """tf.nn"""
  dropout_source = None      # callable(shape, keep_prob) -> uniform [0,1) draws; set by the script

  @staticmethod
  def moments(x, axes, keep_dims=False):
    # nn_impl.moments: mean, then mean of squared_difference(x, stop_gradient(mean)) (population variance)
    a = tuple(axes)
    mean = x.mean(dim=a, keepdim=True)
    var = ((x - mean.detach()) ** 2).mean(dim=a, keepdim=True)
    if not keep_dims:
      for d in sorted(a, reverse=True):
        mean, var = mean.squeeze(d), var.squeeze(d)
    return mean, var

  @staticmethod
  def softmax(x, dim=-1):
    return torch.softmax(x, dim=dim)

  @staticmethod
  def dropout(x, keep_prob):
    # nn_ops.dropout: binary = floor(keep_prob + uniform[0,1)); ret = x / keep_prob * binary
    assert _NN.dropout_source is not None, "set tf.nn.dropout_source (random draws are inputs here)"
    u = _NN.dropout_source(tuple(x.shape), keep_prob).to(x.dtype)
    return x / keep_prob * torch.floor(keep_prob + u)


nn = _NN


class _Image:
  """tf.image -- tensorflow/core/kernels/colorspace_op.h functors (RGBToHSV / HSVToRGB)."""

  @staticmethod
  def rgb_to_hsv(images):
    r, g, b = images[..., 0], images[..., 1], images[..., 2]
    v = images.amax(dim=-1)
    rng = v - images.amin(dim=-1)
    safe_v = torch.where(v > 0, v, torch.ones_like(v))
    s = torch.where(v > 0, rng / safe_v, torch.zeros_like(v))
    safe_r = torch.where(rng > 0, rng, torch.ones_like(rng))
    norm = 1.0 / (6.0 * safe_r)
    hr = norm * (g - b)
    hg = norm * (b - r) + 2.0 / 6.0
    hb = norm * (r - g) + 4.0 / 6.0
    h = torch.where(r == v, hr, torch.where(g == v, hg, hb))
    h = torch.where(rng > 0, h, torch.zeros_like(h))
    h = torch.where(h < 0, h + 1.0, h)
    return torch.stack([h, s, v], dim=-1)

  @staticmethod
  def hsv_to_rgb(images):
    h, s, v = images[..., 0], images[..., 1], images[..., 2]
    dh = h * 6.0
    dr = (torch.abs(dh - 3.0) - 1.0).clamp(0.0, 1.0)
    dg = (-torch.abs(dh - 2.0) + 2.0).clamp(0.0, 1.0)
    db = (-torch.abs(dh - 4.0) + 2.0).clamp(0.0, 1.0)
    one_s = -s + 1.0
    return torch.stack([(one_s + s * dr) * v, (one_s + s * dg) * v, (one_s + s * db) * v], dim=-1)


image = _Image


class _EMA:
  """tf.train.ExponentialMovingAverage(zero_debias=True): after its first update the debiased average IS the
  value (0.01 x / (1 - 0.99)); only the visualisation reads it (net.py:167-168)."""

  def __init__(self, decay, zero_debias=False):
    self.decay = decay

  def apply(self, values):
    return ("ema_update", values)

  def average(self, value):
    return value.detach()


class _Train:
  ExponentialMovingAverage = _EMA

  @staticmethod
  def AdamOptimizer(*a, **k):
    return ("adam", a, k)                    # never stepped here: ly.optimize_loss only records the gradients

  @staticmethod
  def Saver(*a, **k):
    return _Obj()


train = _Train

from . import contrib  # noqa: E402,F401
