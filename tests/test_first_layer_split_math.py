"""The algebra behind csrc/conv_first.cu, on the CPU in float64: the first 4x4 / stride-2 SAME convolution over
concat(image, tile(vec)) - shift equals a convolution over the 3 image channels plus a per-image table indexed by the
output pixel's border class; its weight gradient and the per-image pixel sums of the constant channels' input gradient
follow from class sums of dy.  (The CUDA kernels are checked against autograd in tests/test_nn_gpu.py; this test pins
the derivation itself -- agent.py:11-22, critics.py:51-87, util.py:31-36.)"""
import pytest
import torch
import torch.nn.functional as Fn


def _border_class(o, n):
  return (1 if o == 0 else 0) | (2 if o == n - 1 else 0)


def _tap_inside(k, cls):
  return not ((k == 0 and cls & 1) or (k == 3 and cls & 2))


def _full_conv(x, vec, W, b, shift):
  B, IH, IW, _ = x.shape
  full = torch.cat([x, vec[:, None, None, :].expand(B, IH, IW, vec.shape[1])], dim=3) - shift
  y = Fn.conv2d(full.permute(0, 3, 1, 2), W.permute(3, 2, 0, 1), b, stride=2, padding=1)      # SAME for even sizes
  return y.permute(0, 2, 3, 1), full


@pytest.mark.parametrize("shape", [(2, 8, 8, 5), (1, 2, 2, 3), (2, 4, 6, 11), (1, 16, 2, 14)])
def test_split_forward_wgrad_dgrad(shape):
  B, IH, IW, Cv = shape
  g = torch.Generator().manual_seed(IH * 100 + IW * 10 + Cv)
  R = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
  x, vec = R(B, IH, IW, 3), R(B, Cv)
  W = R(4, 4, 3 + Cv, 32).requires_grad_(True)
  b, shift = R(32), 0.5
  OH, OW = IH // 2, IW // 2
  y_ref, full = _full_conv(x, vec, W, b, shift)
  full = full.detach().requires_grad_(True)
  y_ref = Fn.conv2d(full.permute(0, 3, 1, 2), W.permute(3, 2, 0, 1), b, stride=2, padding=1).permute(0, 2, 3, 1)

  # ---- forward: image channels + border-class table ----
  Wd = W.detach()
  y_img = Fn.conv2d((x - shift).permute(0, 3, 1, 2), Wd[:, :, :3].permute(3, 2, 0, 1), None, stride=2, padding=1).permute(0, 2, 3, 1)
  V = torch.einsum("bc,ykco->byko", vec - shift, Wd[:, :, 3:])                   # [B, ky, kx, co]
  y_split = torch.empty_like(y_img)
  for oy in range(OH):
    for ox in range(OW):
      rc, cc = _border_class(oy, OH), _border_class(ox, OW)
      T = b + sum(V[:, ky, kx] for ky in range(4) for kx in range(4) if _tap_inside(ky, rc) and _tap_inside(kx, cc))
      y_split[:, oy, ox] = y_img[:, oy, ox] + T
  assert torch.allclose(y_split, y_ref.detach(), rtol=1e-12, atol=1e-12)

  # ---- backward from a random dy ----
  dy = R(B, OH, OW, 32)
  gW_ref, gfull = torch.autograd.grad(y_ref, [W, full], grad_outputs=dy)
  # E[b][ky][kx][co] = sum of dy over the output pixels for which the tap falls inside the image
  E = torch.zeros(B, 4, 4, 32, dtype=torch.float64)
  for oy in range(OH):
    for ox in range(OW):
      rc, cc = _border_class(oy, OH), _border_class(ox, OW)
      for ky in range(4):
        for kx in range(4):
          if _tap_inside(ky, rc) and _tap_inside(kx, cc):
            E[:, ky, kx] += dy[:, oy, ox]
  # weight gradient of the constant channels, and of the image channels from the 3-channel convolution alone
  gW_const = torch.einsum("bc,byko->ykco", vec - shift, E)
  assert torch.allclose(gW_const, gW_ref[:, :, 3:], rtol=1e-11, atol=1e-11)
  Wi = Wd[:, :, :3].clone().requires_grad_(True)
  y_i = Fn.conv2d((x - shift).permute(0, 3, 1, 2), Wi.permute(3, 2, 0, 1), None, stride=2, padding=1).permute(0, 2, 3, 1)
  (gW_img,) = torch.autograd.grad(y_i, [Wi], grad_outputs=dy)
  assert torch.allclose(gW_img, gW_ref[:, :, :3], rtol=1e-11, atol=1e-11)
  # input gradient: per-image pixel sums of the constant channels = <W[tap][3 + c], E[b][tap]>
  gvec = torch.einsum("ykco,byko->bc", Wd[:, :, 3:], E)
  assert torch.allclose(gvec, gfull[..., 3:].sum(dim=(1, 2)), rtol=1e-11, atol=1e-11)
