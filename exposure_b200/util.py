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
  """Attribute dict used for cfg and replay records (util.py:40-72)."""

  def __init__(self, *args, **kwargs):
    super(Dict, self).__init__(*args, **kwargs)
    for arg in args:
      if isinstance(arg, dict):
        for k, v in arg.items():
          self[k] = v
    for k, v in kwargs.items():
      self[k] = v

  def __getattr__(self, attr):
    try:
      return self[attr]
    except KeyError:
      raise AttributeError(attr)

  def __setattr__(self, key, value):
    self.__setitem__(key, value)

  def __setitem__(self, key, value):
    super(Dict, self).__setitem__(key, value)
    self.__dict__.update({key: value})

  def __delattr__(self, item):
    self.__delitem__(item)

  def __delitem__(self, key):
    super(Dict, self).__delitem__(key)
    del self.__dict__[key]


def lrelu(x, leak=0.2):
  """util.py:225-229: 0.6 x + 0.4 |x| (derivative 0.6 at 0)."""
  f1 = 0.5 * (1 + leak)
  f2 = 0.5 * (1 - leak)
  return f1 * x + f2 * torch.abs(x)


def rgb2lum(image):
  """util.py:271-274."""
  lum = 0.27 * image[:, :, :, 0] + 0.67 * image[:, :, :, 1] + 0.06 * image[:, :, :, 2]
  return lum[:, :, :, None]


def tanh01(x):
  return torch.tanh(x) * 0.5 + 0.5


def tanh_range(l, r, initial=None):
  """util.py:281-294."""

  def get_activation(left, right, initial):

    def activation(x):
      if initial is not None:
        bias = math.atanh(2 * (initial - left) / (right - left) - 1)
      else:
        bias = 0
      return tanh01(x + bias) * (right - left) + left

    return activation

  return get_activation(l, r, initial)


def lerp(a, b, l):
  """util.py:307-308."""
  return (1 - l) * a + l * b


def enrich_image_input(cfg, net, states):
  """util.py:31-36: tile the state vector over the image and concatenate as channels."""
  if cfg.img_include_states:
    B, H, W, _ = net.shape
    net = torch.cat([net, states[:, None, None, :].expand(B, H, W, states.shape[1])], dim=3)
  return net
