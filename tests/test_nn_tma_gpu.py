"""GPU parity of the TMA-fed tcgen05 convolution engine (backend 4: im2col boxes loaded by
cp.async.bulk.tensor, MN-major TF32 operands in the 128B_BASE32B layout) against the fp64 oracle
and the exact-fp32 CUDA-core engine, over the layer shapes of the policy / value / critic stacks
(agent.py:12-41) plus ragged batches and tiny images that exercise the out-of-bounds fill."""
import pytest
import torch

from oracle import nets as N

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TOL = 2e-5


@pytest.fixture()
def nn(built_lib):
  assert torch.cuda.is_available()
  from exposure_b200 import nn_ops
  nn_ops.set_gemm_backend(nn_ops.BACKEND_TCGEN05_TMA)
  yield nn_ops
  nn_ops.set_gemm_backend(nn_ops.BACKEND_AUTO)


def _rand(*shape, seed=0, scale=1.0):
  return torch.randn(*shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float64) * scale


def _close(a, ref, tol=TOL):
  a = a.detach().cpu().double()
  scale = float(ref.abs().max()) + 1e-30
  err = float((a - ref).abs().max())
  assert err <= tol * scale, "max err %.3g vs scale %.3g (rel %.3g)" % (err, scale, err / scale)


# (B, IH, Cin, Cout): the three deep layers at a ragged batch, a 130-image batch (two M tiles per
# 64-pixel image group), tiny images (tile spans many images), non-power-of-two channel counts
SHAPES = [(3, 32, 32, 64), (5, 16, 64, 128), (7, 8, 128, 256), (130, 8, 32, 32), (9, 4, 32, 64), (33, 2, 64, 32),
          (2, 32, 96, 160), (1, 64, 32, 32)]


@pytest.mark.parametrize("shape", SHAPES)
def test_tma_conv_fprop_dgrad_wgrad(nn, shape):
  B, IH, Cin, Cout = shape
  x = _rand(B, IH, IH, Cin, seed=1) * 0.3
  W = _rand(4, 4, Cin, Cout, seed=2, scale=0.05).requires_grad_(True)
  b = _rand(Cout, seed=3, scale=0.1)
  xin = x.clone().requires_grad_(True)
  pre = N.conv4x4s2(xin, W)
  gy = _rand(*pre.shape, seed=4)
  gin, gW = torch.autograd.grad(pre, [xin, W], grad_outputs=gy)
  f32 = lambda t: t.detach().float().cuda().contiguous()
  # forward (bias + lrelu), dropout side output, tangent mode
  y = nn.conv_fwd(f32(x), f32(W), f32(b))
  _close(y, N.lrelu(pre.detach() + b))
  pm = (torch.rand(pre.shape, generator=torch.Generator().manual_seed(9)) < 0.5).float() * 2
  y1, y2 = nn.conv_fwd(f32(x), f32(W), f32(b), post_mul=pm.cuda())
  assert torch.equal(y1, y) and torch.equal(y2, y * pm.cuda())
  mask = torch.where(y.double().cpu() > 0, 1.0, torch.where(y.double().cpu() < 0, 0.2, 0.6))
  _close(nn.conv_fwd(f32(x), f32(W), None, shift=0.0, mask_ref=y), pre.detach() * mask)
  # dgrad with and without the fused lrelu' of the layer input
  _close(nn.conv_dgrad(f32(gy), f32(W), (B, IH, IH, Cin)), gin)
  a_in = _rand(B, IH, IH, Cin, seed=5)
  a_in[0, 0, 0, :4] = 0.0
  d_mask = torch.where(a_in > 0, 1.0, torch.where(a_in < 0, 0.2, 0.6))
  _close(nn.conv_dgrad(f32(gy), f32(W), (B, IH, IH, Cin), a_in=f32(a_in)), gin * d_mask)
  # wgrad, overwrite and accumulate
  g = nn.conv_wgrad(f32(x), f32(gy))
  _close(g, gW)
  g2 = nn.conv_wgrad(f32(x), f32(gy), out=g.clone(), accumulate=True)
  _close(g2, 2 * gW)
  # and against the CUDA-core engine (same inputs, different summation order)
  nn.set_gemm_backend(nn.BACKEND_CUDA_CORES)
  ys = nn.conv_fwd(f32(x), f32(W), f32(b))
  gs = nn.conv_wgrad(f32(x), f32(gy))
  nn.set_gemm_backend(nn.BACKEND_TCGEN05_TMA)
  _close(y, ys.double().cpu(), tol=3e-5)
  _close(g, gs.double().cpu(), tol=3e-5)


@pytest.mark.parametrize("case", [(3, 64, 3, 3, 32), (5, 64, 3, 11, 32), (2, 64, 3, 14, 32), (2, 32, 3, 11, 64), (67, 8, 5, 4, 32)])
def test_tma_first_layer_staged(nn, case):
  """First layers (Cin = 6 / 14 / 17: image + per-image state constants, shift 0.5) run on the tensor
  cores through a staging copy (zero-bordered 16 channels, or 32 channels for Cin > 16) and padded
  weights; forward, dropout side output, tangent mode, wgrad (overwrite / accumulate)."""
  B, IH, Cx, Cv, Cout = case
  x = _rand(B, IH, IH, Cx, seed=1).abs() * 0.3
  vec = _rand(B, Cv, seed=2)
  W = _rand(4, 4, Cx + Cv, Cout, seed=3, scale=0.1).requires_grad_(True)
  b = _rand(Cout, seed=4, scale=0.1)
  f32 = lambda t: t.detach().float().cuda().contiguous()
  xin = N.enrich(x, vec) - 0.5
  pre = N.conv4x4s2(xin, W)
  gy = _rand(*pre.shape, seed=5)
  (gW,) = torch.autograd.grad(pre, [W], grad_outputs=gy)
  y = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=0.5)
  _close(y, N.lrelu(pre.detach() + b))
  pm = (torch.rand(pre.shape, generator=torch.Generator().manual_seed(9)) < 0.5).float() * 2
  y1, y2 = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=0.5, post_mul=pm.cuda())
  assert torch.equal(y1, y) and torch.equal(y2, y * pm.cuda())
  t_img, t_vec = _rand(B, IH, IH, Cx, seed=7), _rand(B, Cv, seed=8)
  mask = torch.where(y.double().cpu() > 0, 1.0, torch.where(y.double().cpu() < 0, 0.2, 0.6))
  _close(nn.conv_fwd(f32(t_img), f32(W), None, vec=f32(t_vec), shift=0.0, mask_ref=y),
         N.conv4x4s2(N.enrich(t_img, t_vec), W.detach()) * mask)
  g = nn.conv_wgrad(f32(x), f32(gy), vec=f32(vec), shift=0.5)
  _close(g, gW)
  g2 = nn.conv_wgrad(f32(x), f32(gy), vec=f32(vec), shift=0.5, out=g.clone(), accumulate=True)
  _close(g2, 2 * gW)
  nn.set_gemm_backend(nn.BACKEND_CUDA_CORES)
  ys = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=0.5)
  gs = nn.conv_wgrad(f32(x), f32(gy), vec=f32(vec), shift=0.5)
  nn.set_gemm_backend(nn.BACKEND_TCGEN05_TMA)
  _close(y, ys.double().cpu(), tol=3e-5)
  _close(g, gs.double().cpu(), tol=3e-5)


@pytest.mark.parametrize("case", [(3, 64, 6, 32), (2, 64, 17, 32), (5, 16, 14, 32), (2, 8, 3, 64), (66, 4, 20, 32)])
def test_dgrad_into_few_channels(nn, case):
  """Layer-1 dgrad (gradient w.r.t. the enriched image, Cin = 6 critic / 17 value network) runs on the
  dedicated thread-per-pixel kernel under the AUTO / TMA backends; exact fp32."""
  B, IH, Cin, Cout = case
  W = _rand(4, 4, Cin, Cout, seed=2, scale=0.05)
  xin = _rand(B, IH, IH, Cin, seed=1).requires_grad_(True)
  pre = N.conv4x4s2(xin, W)
  gy = _rand(*pre.shape, seed=4)
  (gin,) = torch.autograd.grad(pre, [xin], grad_outputs=gy)
  f32 = lambda t: t.detach().float().cuda().contiguous()
  _close(nn.conv_dgrad(f32(gy), f32(W), (B, IH, IH, Cin)), gin, tol=2e-6)
  a_in = _rand(B, IH, IH, Cin, seed=5)
  d_mask = torch.where(a_in > 0, 1.0, torch.where(a_in < 0, 0.2, 0.6))
  _close(nn.conv_dgrad(f32(gy), f32(W), (B, IH, IH, Cin), a_in=f32(a_in)), gin * d_mask, tol=2e-6)
