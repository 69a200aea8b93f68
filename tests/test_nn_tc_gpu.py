"""GPU parity of the tcgen05 (tensor-core, 3xTF32) instantiations of the network primitives
against the fp64 oracle and against the exact-fp32 CUDA-core engine.  Tolerance 2e-5 * max|ref|
(3xTF32 keeps ~2^-22 relative accuracy; plain TF32 would be ~5e-4 and fail this test)."""
import pytest
import torch

from oracle import nets as N

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TOL = 2e-5


@pytest.fixture(params=[2, 3, 4], ids=["tcgen05", "tcgen05-warp-specialised", "tcgen05-tma"])
def nn(built_lib, request):
  assert torch.cuda.is_available()
  from exposure_b200 import nn_ops
  nn_ops.BACKEND_UNDER_TEST = request.param
  nn_ops.set_gemm_backend(request.param)
  yield nn_ops
  nn_ops.set_gemm_backend(nn_ops.BACKEND_AUTO)


def _rand(*shape, seed=0, scale=1.0):
  return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


def _close(a, ref, tol=TOL):
  a = a.detach().cpu().double()
  scale = float(ref.abs().max()) + 1e-30
  err = float((a - ref).abs().max())
  assert err <= tol * scale, "max err %.3g vs scale %.3g (rel %.3g)" % (err, scale, err / scale)


CASES = [(3, 32, 32, 0, 64, 0.0), (2, 16, 64, 0, 128, 0.0), (5, 8, 128, 0, 256, 0.0), (3, 64, 3, 11, 32, 0.5),
         (2, 64, 3, 3, 32, 0.5), (2, 64, 3, 14, 32, 0.5), (1, 2, 32, 0, 32, 0.0)]


@pytest.mark.parametrize("case", CASES)
def test_tc_conv_forward_and_tangent(nn, case):
  B, IH, Cx, Cv, Cout, shift = case
  x = _rand(B, IH, IH, Cx, seed=1).abs() * 0.3
  vec = _rand(B, Cv, seed=2) if Cv else None
  W = _rand(4, 4, Cx + Cv, Cout, seed=3, scale=0.1)
  b = _rand(Cout, seed=4, scale=0.1)
  xin = N.enrich(x, vec) if Cv else x
  y = N.lrelu(N.conv4x4s2(xin - shift, W, b))
  f32 = lambda t: None if t is None else t.float().cuda().contiguous()
  yd = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=shift)
  _close(yd, y)
  nn.set_gemm_backend(nn.BACKEND_CUDA_CORES)
  ys = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=shift)
  nn.set_gemm_backend(nn.BACKEND_UNDER_TEST)
  _close(yd, ys.double().cpu(), tol=3e-5)    # two different fp32 summation orders over K up to 2048
  pm = (torch.rand(y.shape, generator=torch.Generator().manual_seed(9)) < 0.5).float() * 2
  y1, y2 = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=shift, post_mul=pm.cuda())
  assert torch.equal(y1, yd) and torch.equal(y2, yd * pm.cuda())
  t_in = _rand(B, IH, IH, Cx, seed=7)
  tvec = _rand(B, Cv, seed=8) if Cv else None
  tin_full = N.enrich(t_in, tvec) if Cv else t_in
  mask = torch.where(ys.double().cpu() > 0, 1.0, torch.where(ys.double().cpu() < 0, 0.2, 0.6))
  t_ref = N.conv4x4s2(tin_full, W) * mask
  t_out = nn.conv_fwd(f32(t_in), f32(W), None, vec=f32(tvec), shift=0.0, mask_ref=ys)
  _close(t_out, t_ref)


@pytest.mark.parametrize("case", [(64, 4096, 128), (192, 4096, 128), (64, 4096, 1024), (64, 128, 30), (7, 100, 16)])
def test_tc_fc_forward(nn, case):
  M, K, Nn = case
  x = _rand(M, K, seed=1)
  W = _rand(K, Nn, seed=2, scale=K ** -0.5)
  b = _rand(Nn, seed=3, scale=0.1)
  f32 = lambda t: t.float().cuda().contiguous()
  pre = x @ W + b
  _close(nn.fc_fwd(f32(x), f32(W), f32(b), mode=nn.FC_LRELU), N.lrelu(pre))
  _close(nn.fc_fwd(f32(x), f32(W), f32(b), mode=nn.FC_LINEAR), pre)
  big = torch.zeros(M, 2 * K)
  big[:, K:] = x.float()
  _close(nn.fc_fwd(big.cuda()[:, K:], f32(W), f32(b), mode=nn.FC_LINEAR), pre)


def test_tc_backward_primitives(nn):
  """conv dgrad / wgrad and FC dgrad / wgrad on the tensor-core engines vs the fp64 oracle."""
  B, IH, Cin, Cout = 3, 16, 64, 128
  x = _rand(B, IH, IH, Cin, seed=1) * 0.3
  W = _rand(4, 4, Cin, Cout, seed=2, scale=0.05).requires_grad_(True)
  xin = x.clone().requires_grad_(True)
  y = N.conv4x4s2(xin, W)
  gy = _rand(*y.shape, seed=3)
  gin, gW = torch.autograd.grad(y, [xin, W], grad_outputs=gy)
  f32 = lambda t: t.detach().float().cuda().contiguous()
  _close(nn.conv_dgrad(f32(gy), f32(W), (B, IH, IH, Cin)), gin)
  _close(nn.conv_wgrad(f32(x), f32(gy)), gW)
  M, K, Nn = 96, 4096, 128
  a = _rand(M, K, seed=4).requires_grad_(True)
  Wf = _rand(K, Nn, seed=5, scale=K ** -0.5).requires_grad_(True)
  out = a @ Wf
  go = _rand(M, Nn, seed=6)
  ga, gWf = torch.autograd.grad(out, [a, Wf], grad_outputs=go)
  _close(nn.fc_dgrad(f32(go), f32(Wf)), ga)
  _close(nn.fc_wgrad(f32(a), f32(go)), gWf)
