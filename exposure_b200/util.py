"""Host-side helpers mirroring the reference's util.py names the hot path uses
(util.py:13-16, 31-36, 40-72, 225-229, 271-308).  Tensors are torch CUDA tensors; these
helpers are API glue (per-image / tiny), the per-pixel work lives in csrc/."""
import math

import torch

STATE_REWARD_DIM = 0       # util.py:13-16
STATE_STOPPED_DIM = 1
STATE_STEP_DIM = 2
STATE_DROPOUT_BEGIN = 3


class Dict(dict):
  """Attribute-style dict for cfg and replay records (same behaviour as util.py:40-72: keys are readable,
  writable and deletable as attributes; nested dict arguments are merged in)."""

  def __init__(self, *sources, **items):
    dict.__init__(self)
    for src in sources:
      self.update(src)
    self.update(items)

  def __getattr__(self, name):
    if name in self:
      return self[name]
    raise AttributeError(name)

  def __setattr__(self, name, value):
    self[name] = value

  def __delattr__(self, name):
    if name not in self:
      raise AttributeError(name)
    del self[name]


def lrelu(x, leak=0.2):
  """util.py:225-229: leaky relu written as 0.5(1+leak) x + 0.5(1-leak) |x| (derivative 0.6 at 0 for leak 0.2)."""
  return (0.5 * (1 + leak)) * x + (0.5 * (1 - leak)) * torch.abs(x)


_LUM_WEIGHTS = (0.27, 0.67, 0.06)          # util.py:271-274


def rgb2lum(image):
  """Luminance of an NHWC batch, kept as a trailing singleton channel (util.py:271-274)."""
  r, g, b = image[..., 0], image[..., 1], image[..., 2]
  return (_LUM_WEIGHTS[0] * r + _LUM_WEIGHTS[1] * g + _LUM_WEIGHTS[2] * b).unsqueeze(-1)


def tanh01(x):
  return 0.5 * torch.tanh(x) + 0.5


def tanh_range(l, r, initial=None):
  """util.py:281-294: x -> l + (r - l) tanh01(x + bias), with bias chosen so that x = 0 maps to `initial`."""
  span = r - l
  bias = 0 if initial is None else math.atanh(2 * (initial - l) / span - 1)
  return lambda x: tanh01(x + bias) * span + l


def lerp(a, b, l):
  """util.py:307-308."""
  return a * (1 - l) + b * l


def enrich_image_input(cfg, net, states):
  """util.py:31-36: tile the state vector over the image and concatenate as channels."""
  if not cfg.img_include_states:
    return net
  B, H, W, _ = net.shape
  tiled = states[:, None, None, :].expand(B, H, W, states.shape[1])
  return torch.cat([net, tiled], dim=3)
