"""GPU parity of the network primitives (conv / FC, through the C ABI) against the CPU oracle
(oracle/nets.py, torch CPU fp64 as arbiter).  These kernels accumulate in exact fp32 (no TF32),
so the tolerance is fp32 summation noise: |err| <= 2e-5 * max|ref| (north_star: logits <= 1e-3 rel)."""
import pytest
import torch

from oracle import nets as N

pytestmark = pytest.mark.gpu
TOL = 2e-5


@pytest.fixture(scope="module")
def nn(built_lib):
  assert torch.cuda.is_available()
  from exposure_b200 import nn_ops
  return nn_ops


def _close(a, ref, tol=TOL):
  a = a.detach().cpu().double()
  scale = float(ref.abs().max()) + 1e-30
  err = float((a - ref).abs().max())
  assert err <= tol * scale, "max err %.3g vs scale %.3g" % (err, scale)


def _rand(*shape, seed=0, scale=1.0):
  return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


CONV_CASES = [  # B, IH, Cx, Cv, Cout, shift
    (3, 64, 3, 11, 32, 0.5),     # policy layer 1 (agent.py:21-22), enriched input
    (2, 64, 3, 3, 32, 0.5),      # critic layer 1 (critics.py:13-19)
    (2, 64, 3, 14, 32, 0.5),     # value layer 1 (17 channels)
    (3, 32, 32, 0, 64, 0.0),
    (3, 16, 64, 0, 128, 0.0),
    (5, 8, 128, 0, 256, 0.0),
    (1, 2, 8, 0, 8, 0.0),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_forward_backward(nn, case):
  B, IH, Cx, Cv, Cout, shift = case
  x = _rand(B, IH, IH, Cx, seed=1).abs() * 0.3
  vec = _rand(B, Cv, seed=2) if Cv else None
  W = _rand(4, 4, Cx + Cv, Cout, seed=3, scale=0.1).requires_grad_(True)
  b = _rand(Cout, seed=4, scale=0.1).requires_grad_(True)
  xin = (N.enrich(x, vec) if Cv else x).clone().requires_grad_(True)
  pre = N.conv4x4s2(xin - shift, W, b)
  y = N.lrelu(pre)
  f32 = lambda t: None if t is None else t.detach().float().cuda().contiguous()
  yd = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=shift)
  _close(yd, y.detach())
  # dropout-style second output
  pm = (torch.rand(y.shape, generator=torch.Generator().manual_seed(9)) < 0.5).float() * 2
  y1, y2 = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=shift, post_mul=pm.cuda())
  assert torch.equal(y1, yd) and torch.equal(y2, yd * pm.cuda())
  # backward: delta = dL/dpre given dL/dy ; dgrad, wgrad, bias grad
  gy = _rand(*y.shape, seed=5)
  gin, gW, gb = torch.autograd.grad(y, [xin, W, b], grad_outputs=gy)
  delta = (gy * torch.where(pre > 0, 1.0, torch.where(pre < 0, 0.2, 0.6))).detach()
  dx = nn.conv_dgrad(f32(delta), f32(W), (B, IH, IH, Cx + Cv))
  _close(dx, gin)
  gWd = nn.conv_wgrad(f32(x), f32(delta), vec=f32(vec), shift=shift)
  _close(gWd, gW)
  _close(nn.colsum(f32(delta)), gb)
  # dgrad fused with the previous layer's lrelu derivative
  a_in = _rand(B, IH, IH, Cx + Cv, seed=6)
  a_in[0, 0, 0, 0] = 0.0
  dmask = torch.where(a_in > 0, 1.0, torch.where(a_in < 0, 0.2, 0.6))
  dx2 = nn.conv_dgrad(f32(delta), f32(W), (B, IH, IH, Cx + Cv), a_in=f32(a_in))
  _close(dx2, gin * dmask)
  # forward-mode tangent: t_out = conv_nobias(t_in) * lrelu'(a_out)
  t_in = _rand(B, IH, IH, Cx, seed=7)
  tvec = _rand(B, Cv, seed=8) if Cv else None
  tin_full = N.enrich(t_in, tvec) if Cv else t_in
  t_ref = N.conv4x4s2(tin_full, W.detach()) * torch.where(y.detach() > 0, 1.0, torch.where(y.detach() < 0, 0.2, 0.6))
  t_out = nn.conv_fwd(f32(t_in), f32(W), None, vec=f32(tvec), shift=0.0, mask_ref=yd)
  _close(t_out, t_ref, tol=1e-4)   # mask taken from the fp32 activation: sign flips of ~0 values allowed for


FC_CASES = [(64, 4096, 128), (192, 4096, 128), (64, 128, 8), (64, 128, 30), (64, 128, 1), (7, 100, 9)]


@pytest.mark.parametrize("case", FC_CASES)
def test_fc_forward_backward(nn, case):
  M, K, Nn = case
  x = _rand(M, K, seed=1).requires_grad_(True)
  W = _rand(K, Nn, seed=2, scale=K ** -0.5).requires_grad_(True)
  b = _rand(Nn, seed=3, scale=0.1).requires_grad_(True)
  f32 = lambda t: t.detach().float().cuda().contiguous()
  pre = x @ W + b
  y = N.lrelu(pre)
  yd = nn.fc_fwd(f32(x), f32(W), f32(b), mode=nn.FC_LRELU)
  _close(yd, y.detach())
  _close(nn.fc_fwd(f32(x), f32(W), f32(b), mode=nn.FC_LINEAR), pre.detach())
  gy = _rand(M, Nn, seed=4)
  gx, gW, gb = torch.autograd.grad(y, [x, W, b], grad_outputs=gy)
  delta = (gy * torch.where(pre > 0, 1.0, torch.where(pre < 0, 0.2, 0.6))).detach()
  _close(nn.fc_dgrad(f32(delta), f32(W)), gx)
  _close(nn.fc_wgrad(f32(x), f32(delta)), gW)
  _close(nn.colsum(f32(delta)), gb)
  mul = _rand(M, K, seed=5)
  _close(nn.fc_dgrad(f32(delta), f32(W), mul=f32(mul), mul_mode=2), gx * mul)
  _close(nn.fc_dgrad(f32(delta), f32(W), mul=f32(mul), mul_mode=1), gx * torch.where(mul > 0, 1.0, 0.2))
  # tangent
  t = _rand(M, K, seed=6)
  t_ref = (t @ W.detach()) * torch.where(y.detach() > 0, 1.0, torch.where(y.detach() < 0, 0.2, 0.6))
  _close(nn.fc_fwd(f32(t), f32(W), None, mode=nn.FC_TANGENT, mask_ref=yd), t_ref, tol=1e-4)
