"""tf.contrib stand-in (see ../__init__.py): contrib.layers and the one distribution net.py:175 samples."""
import torch

from . import layers  # noqa: F401


class _Uniform:
  """tf.contrib.distributions.Uniform(low, high).sample(shape): draws come from `distributions.source`
  (callable(shape) -> float64 tensor in [0,1)) and are logged so they can be replayed elsewhere."""

  def __init__(self, low=0.0, high=1.0):
    self.low, self.high = low, high

  def sample(self, shape):
    import tensorflow as tf
    assert distributions.source is not None, "set tf.contrib.distributions.source (random draws are inputs here)"
    u = distributions.source(tuple(int(s) for s in shape))
    distributions.log.append(u)
    return (self.low + (self.high - self.low) * u).to(tf._state["float"])


class distributions:
  Uniform = _Uniform
  source = None
  log = []
