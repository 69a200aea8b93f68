"""Mapping between the trainer's flat parameter stores and the reference checkpoint's variable
names / layouts: conv kernels HWIO, FC [in,out], NHWC flatten.

The only layout difference is internal: the 8 `generator/filter_j/fc1` layers are stored as one
[4096, 8*128] matrix (column block j) so they run as a single GEMM."""
import torch

from .nets import FC1


def export_named(trainer, grads=False):
  """dict: reference variable name -> tensor (copies), for theta_g, theta_v, theta_c."""
  out = {"generator": {}, "rl_value": {}, "critic": {}}
  for key, store in (("generator", trainer.gen), ("rl_value", trainer.val), ("critic", trainer.cri)):
    src = store.g if grads else store.p
    for name, t in src.items():
      if "/filter_fc1_all/" in name:
        kind = name.rsplit("/", 1)[1]
        for j in range(8):
          blk = t[:, j * FC1:(j + 1) * FC1] if kind == "weights" else t[j * FC1:(j + 1) * FC1]
          out[key]["generator/filter_%d/fc1/%s" % (j, kind)] = blk.detach().clone()
      else:
        out[key][name] = t.detach().clone()
  return out


def import_named(trainer, named):
  """Inverse of export_named for parameters: `named` maps reference variable names -> tensors."""
  for store in (trainer.gen, trainer.val, trainer.cri):
    for name, t in store.p.items():
      if "/filter_fc1_all/" in name:
        kind = name.rsplit("/", 1)[1]
        for j in range(8):
          src = named["generator/filter_%d/fc1/%s" % (j, kind)].to(t.device, torch.float32)
          if kind == "weights":
            t[:, j * FC1:(j + 1) * FC1].copy_(src)
          else:
            t[j * FC1:(j + 1) * FC1].copy_(src)
      elif name in named:
        t.copy_(named[name].to(t.device, torch.float32).reshape(t.shape))
      else:
        raise KeyError("variable %s missing from the checkpoint" % name)
