"""Data parallelism of the train step: one process per GPU, batch sharded by image.

Every filter parameter and every CNN activation is per image, so the forward / backward passes need no
communication; the only exchange is the gradient of each optimizer step.  All losses are batch means
(net.py:129, 151, 162, 186), so the mean over ranks of the per-shard gradients IS the global-batch gradient.

Two transports behind `Trainer._apply`:

* `PeerExchange` (NCCL process groups on one box, the product path): every rank maps every peer's gradient
  buffer into its address space once (CUDA IPC, handles travel through `dist.all_gather_object`) and the step
  ends in ONE kernel, `exp_dp_allreduce_adam` (csrc/dp.cu): reduce-scatter + all-gather over NVLink peer loads
  fused with Adam, flag barriers instead of host synchronisation -- a plain kernel node of the step's CUDA graph.
* `dist.all_reduce` on the flat buffer followed by the fused Adam with grad_scale = 1/world: any other process
  group (gloo in the tests; a box without peer access).  Not capturable, so these steps replay the graph and
  then run the two calls eagerly."""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _cabi


def world_info():
  if dist.is_available() and dist.is_initialized():
    return dist.get_rank(), dist.get_world_size()
  return 0, 1


def shard_range(n, rank, world):
  """Contiguous, balanced [begin, end) slice of n items for `rank` (sizes differ by <= 1)."""
  base, rem = divmod(n, world)
  begin = rank * base + min(rank, rem)
  return begin, begin + base + (1 if rank < rem else 0)


def rank_seed(base, rank, stream=0):
  """Distinct, reproducible RNG seeds per (rank, stream): dropout, z, alpha, data order."""
  return int(base) * 1000003 + int(rank) * 101 + int(stream)


def peer_exchange_possible(device):
  """The peer-memory transport needs: NCCL-style one-GPU-per-process groups on ONE node, <= exp_dp_max_world()
  ranks, every rank on its own device.  EXPOSURE_DP_TRANSPORT=allreduce forces the fallback (A/B switch)."""
  rank, world = world_info()
  if world < 2 or os.environ.get("EXPOSURE_DP_TRANSPORT", "peer") != "peer":
    return False
  if dist.get_backend() != "nccl" or device.type != "cuda" or world > _cabi.lib().exp_dp_max_world():
    return False
  dev = torch.tensor([device.index], device=device)
  devs = [torch.zeros_like(dev) for _ in range(world)]
  dist.all_gather(devs, dev)
  host = [None] * world
  dist.all_gather_object(host, os.uname().nodename)
  return len({int(d) for d in devs}) == world and len(set(host)) == 1


_opened = {}      # IPC handle bytes -> mapped base address in this process (one open per handle and process)


def _map_peer(handle, offset):
  base = _opened.get(handle)
  if base is None:
    out = ctypes.c_void_p()
    buf = ctypes.create_string_buffer(handle, len(handle))
    _cabi.check(_cabi.lib().exp_dp_ipc_open(ctypes.cast(buf, ctypes.c_void_p), ctypes.byref(out)), "exp_dp_ipc_open")
    base = _opened[handle] = int(out.value)
  return base + int(offset)


class PeerExchange:
  """Peer-mapped gradient exchange of ONE flat gradient buffer (theta_g + theta_v, or theta_c)."""

  def __init__(self, grad, device):
    self.rank, self.world = world_info()
    self.device = device
    self.n = grad.numel()
    if self.n % (4 * self.world):
      raise ValueError("flat buffer of %d floats is not a multiple of 4 * world" % self.n)
    l = _cabi.lib()
    self.grad = grad
    self.red = torch.zeros_like(grad)
    self.flags = torch.zeros(l.exp_dp_flag_bytes() // 4, dtype=torch.int32, device=device)
    torch.cuda.synchronize(device)
    hb = l.exp_dp_ipc_handle_bytes()
    self._ptrs = []
    for t in (self.grad, self.red, self.flags):
      handle = ctypes.create_string_buffer(hb)
      offset = ctypes.c_size_t()
      _cabi.check(l.exp_dp_ipc_export(t.data_ptr(), ctypes.cast(handle, ctypes.c_void_p), ctypes.byref(offset)), "exp_dp_ipc_export")
      gathered = [None] * self.world
      dist.all_gather_object(gathered, (handle.raw, int(offset.value)))
      ptrs = [t.data_ptr() if q == self.rank else _map_peer(*gathered[q]) for q in range(self.world)]
      self._ptrs.append((ctypes.c_void_p * self.world)(*ptrs))
    dist.barrier()                                 # every rank holds every mapping before the first launch

  def allreduce_adam(self, params, m, v, hyper_a, n_a, hyper_b, beta1, beta2, eps=1e-8):
    from . import nn_ops as K
    g, r, f = self._ptrs
    _cabi.check(_cabi.lib().exp_dp_allreduce_adam(
        params.data_ptr(), m.data_ptr(), v.data_ptr(), ctypes.cast(g, ctypes.c_void_p), ctypes.cast(r, ctypes.c_void_p),
        ctypes.cast(f, ctypes.c_void_p), self.world, self.rank, hyper_a.data_ptr(), int(n_a),
        hyper_b.data_ptr() if hyper_b is not None else None, self.n, float(beta1), float(beta2), float(eps),
        torch.cuda.current_stream().cuda_stream), "exp_dp_allreduce_adam")
    K._n()
