"""N-step filter chain driver (BASELINE.json configs[1]/[4]: "8-filter chain fwd+bwd").

The chain x_0 -> f_{id_0} -> x_1 -> ... -> x_N is a benchmark construct built from the
reference's per-filter maths (filters.py process()); in the reference one step applies one
selected filter (agent.py:113-125) and an episode is cfg.test_steps such steps.  Each step is
ONE fused CUDA kernel forward and ONE backward (exposure_b200/csrc/filters.cu); activations
x_0..x_{N-1} stay resident in HBM for the backward (12 B/pixel/step), outputs are recomputed.

Explicit schedule, no autograd: forward() then backward()."""
import torch

from . import ops


class FilterChain:

  def __init__(self, ids, variant=ops.VARIANT_AUTO):
    """ids: list of N entries, each an int filter id (uniform over the batch) or a CUDA int32
    tensor [B] of per-image ids."""
    self.ids = list(ids)
    self.variant = variant
    self._acts = None
    self._params = None
    self._logits = None
    self._gbuf = None

  def _alloc(self, x):
    n = len(self.ids)
    if self._acts is None or self._acts[0].shape != x.shape or self._acts[0].device != x.device:
      self._acts = [torch.empty_like(x) for _ in range(n + 1)]
      self._gbuf = [torch.empty_like(x) for _ in range(2)]

  def forward(self, x, logits_list):
    """x: [B,H,W,3]; logits_list[k]: [B, >= n_k] raw regressor inputs.  Returns x_N."""
    assert len(logits_list) == len(self.ids)
    self._alloc(x)
    self._acts[0].copy_(x) if self._acts[0].data_ptr() != x.data_ptr() else None
    self._logits = [l.contiguous() for l in logits_list]
    self._params = []
    for k, fid in enumerate(self.ids):
      p = ops.filter_regress_fwd(self._logits[k], fid)
      self._params.append(p)
      ops.filter_fwd(self._acts[k], p, fid, out=self._acts[k + 1], variant=self.variant)
    return self._acts[-1]

  def forward_resident(self, logits_list):
    """Same as forward() with x_0 already resident in the chain's first activation buffer
    (see input_buffer()); used by the benchmark's device-resident leg."""
    return self.forward(self._acts[0], logits_list)

  def input_buffer(self, shape, device):
    self._alloc(torch.empty(shape, device=device, dtype=torch.float32))
    return self._acts[0]

  def backward(self, gout, need_input_grad=True):
    """gout = dL/dx_N.  Returns (dL/dx_0 or None, [dL/dlogits_k])."""
    n = len(self.ids)
    g = gout
    glogits = [None] * n
    for k in reversed(range(n)):
      fid = self.ids[k]
      need_gx = need_input_grad or k > 0
      gx, gparams = ops.filter_bwd(self._acts[k], g, self._params[k], fid, need_gx=need_gx,
                                   gx_out=self._gbuf[k & 1] if need_gx else None, variant=self.variant)
      glogits[k] = ops.filter_regress_bwd(self._logits[k], gparams, fid)
      g = gx
    return g, glogits

