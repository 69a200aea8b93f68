"""tf.contrib.layers stand-in (see ../__init__.py): conv2d / fully_connected / dropout with the
variable naming of TF 1.x (scope default names "Conv", "Conv_1", ..., "fully_connected",
"fully_connected_1", variables "weights" and "biases"), NHWC activations, HWIO kernels, SAME padding."""
import torch
import torch.nn.functional as F

import tensorflow as tf


def xavier_initializer(*a, **k):
  return "xavier"


def _relu(x):
  return torch.relu(x)


def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn=_relu, normalizer_fn=None,
           normalizer_params=None, weights_initializer=None, scope=None, reuse=None):
  assert normalizer_fn is None and padding == "SAME"
  cout, k, s = int(num_outputs), int(kernel_size), int(stride)
  cin = int(inputs.shape[3])
  with tf.variable_scope(scope, default_name="Conv"):
    w = tf.get_variable("weights", (k, k, cin, cout))
    b = tf.get_variable("biases", (cout,))
  h, wd = int(inputs.shape[1]), int(inputs.shape[2])
  # SAME: out = ceil(in / s); total pad = max((out-1)*s + k - in, 0), the odd one goes to the end
  def pads(n):
    out = -(-n // s)
    tot = max((out - 1) * s + k - n, 0)
    return tot // 2, tot - tot // 2
  (pt, pb), (pl, pr) = pads(h), pads(wd)
  x = F.pad(inputs.permute(0, 3, 1, 2), (pl, pr, pt, pb))
  y = F.conv2d(x, w.permute(3, 2, 0, 1), bias=b, stride=s).permute(0, 2, 3, 1)
  return activation_fn(y) if activation_fn is not None else y


conv = conv2d


def fully_connected(inputs, num_outputs, activation_fn=_relu, weights_initializer=None, scope=None, reuse=None):
  n_in, n_out = int(inputs.shape[-1]), int(num_outputs)
  with tf.variable_scope(scope, default_name="fully_connected"):
    w = tf.get_variable("weights", (n_in, n_out))
    b = tf.get_variable("biases", (n_out,))
  y = inputs @ w + b
  return activation_fn(y) if activation_fn is not None else y


def dropout(inputs, keep_prob=0.5, is_training=True, **k):
  return tf.nn.dropout(inputs, keep_prob) if is_training else inputs


optimize_log = []      # one record per ly.optimize_loss call: the loss and d loss / d variable by name


def optimize_loss(loss, learning_rate, optimizer, variables, global_step=None, summaries=None, **k):
  """tf.contrib.layers.optimize_loss: here only its gradient computation (tf.gradients(loss, variables));
  nothing is stepped."""
  grads = torch.autograd.grad(loss, [v.tensor for v in variables], retain_graph=True, allow_unused=True)
  rec = {"loss": loss.detach(), "grads": {v.name: (torch.zeros_like(v.tensor) if g is None else g.detach())
                                          for v, g in zip(variables, grads)}}
  optimize_log.append(rec)
  return rec
