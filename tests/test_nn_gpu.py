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
    (2, 16, 6, 0, 32, 0.0),      # small-Cin dgrad: tile narrower than a warp (OW = 8)
    (1, 128, 3, 3, 32, 0.5),     # small-Cin dgrad: two tiles per row (OW = 64)
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
  # the mask comes from the activation the kernel is handed (fp32): an fp64 activation within rounding of 0 can have
  # the other sign, which is a property of the input, not of the kernel
  ym = yd.cpu().double()
  t_ref = N.conv4x4s2(tin_full, W.detach()) * torch.where(ym > 0, 1.0, torch.where(ym < 0, 0.2, 0.6))
  t_out = nn.conv_fwd(f32(t_in), f32(W), None, vec=f32(tvec), shift=0.0, mask_ref=yd)
  _close(t_out, t_ref, tol=1e-4)


@pytest.mark.parametrize("case", [(3, 64, 11), (2, 64, 3), (2, 64, 14), (1, 128, 3), (2, 16, 3), (2, 4, 3), (1, 2, 5), (2, 64, 0)])
def test_first_layer_split(nn, case):
  """exp_conv_first_*: image channels by a direct kernel, per-image constant channels through the border-class table
  (forward, tangent mode, weight gradient, and the input gradient as its producers need it)."""
  B, IH, Cv = case
  Cx, Cout = 3, 32
  x = _rand(B, IH, IH, Cx, seed=11).abs() * 0.3
  vec = _rand(B, Cv, seed=12) if Cv else None
  W = _rand(4, 4, Cx + Cv, Cout, seed=13, scale=0.1)
  b = _rand(Cout, seed=14, scale=0.1)
  xin = (N.enrich(x, vec) if Cv else x).clone().requires_grad_(True)
  Wr = W.clone().requires_grad_(True)
  pre = N.conv4x4s2(xin - 0.5, Wr, b)
  y = N.lrelu(pre)
  f32 = lambda t: None if t is None else t.detach().float().cuda().contiguous()
  assert nn.first_layer_split(Cx + Cv)
  yd = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=0.5)
  _close(yd, y.detach())
  pm = (torch.rand(y.shape, generator=torch.Generator().manual_seed(9)) < 0.5).float() * 2
  y1, y2 = nn.conv_fwd(f32(x), f32(W), f32(b), vec=f32(vec), shift=0.5, post_mul=pm.cuda())
  assert torch.equal(y1, yd) and torch.equal(y2, yd * pm.cuda())
  gy = _rand(*y.shape, seed=15)
  gin, gW = torch.autograd.grad(y, [xin, Wr], grad_outputs=gy)
  delta = (gy * torch.where(pre > 0, 1.0, torch.where(pre < 0, 0.2, 0.6))).detach()
  gWd = nn.conv_wgrad(f32(x), f32(delta), vec=f32(vec), shift=0.5)
  _close(gWd, gW)
  acc = torch.full_like(gWd, 0.25)
  nn.conv_wgrad(f32(x), f32(delta), vec=f32(vec), shift=0.5, out=acc, accumulate=True)
  _close(acc - 0.25, gW, tol=2e-5)
  dx, gv = nn.conv_first_dgrad(f32(delta), f32(W), Cv, (IH, IH))
  _close(dx, gin[..., :Cx])
  if Cv:
    _close(gv, gin[..., Cx:].sum(dim=(1, 2)))
  # bit-for-bit repeatable (fixed summation order)
  assert torch.equal(gWd, nn.conv_wgrad(f32(x), f32(delta), vec=f32(vec), shift=0.5))
  # tangent mode
  t_in = _rand(B, IH, IH, Cx, seed=17)
  tvec = _rand(B, Cv, seed=18) if Cv else None
  ym = yd.cpu().double()
  t_ref = N.conv4x4s2(N.enrich(t_in, tvec) if Cv else t_in, W) * torch.where(ym > 0, 1.0, torch.where(ym < 0, 0.2, 0.6))
  t_out = nn.conv_fwd(f32(t_in), f32(W), None, vec=f32(tvec), shift=0.0, mask_ref=yd)
  _close(t_out, t_ref, tol=1e-5)


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
  _close(nn.fc_dgrad(f32(delta), f32(W), mul_plain=f32(mul)), gx * mul)
  _close(nn.fc_dgrad(f32(delta), f32(W), mul_act=f32(mul)), gx * torch.where(mul > 0, 1.0, 0.2))
  _close(nn.fc_dgrad(f32(delta), f32(W), mul_act=f32(mul), mul_plain=f32(mul)), gx * torch.where(mul > 0, 1.0, 0.2) * mul)
  acc = f32(gx).clone()
  nn.fc_dgrad(f32(delta), f32(W), out=acc, accumulate=True)
  _close(acc, 2 * gx)
  accw = f32(gW).clone()
  nn.fc_wgrad(f32(x), f32(delta), out=accw, accumulate=True)
  _close(accw, 2 * gW)
  # tangent
  t = _rand(M, K, seed=6)
  t_ref = (t @ W.detach()) * torch.where(y.detach() > 0, 1.0, torch.where(y.detach() < 0, 0.2, 0.6))
  _close(nn.fc_fwd(f32(t), f32(W), None, mode=nn.FC_TANGENT, mask_ref=yd), t_ref, tol=1e-4)


def test_fc_column_slices(nn):
  """Leading dimensions: heads that share one activation matrix (8 filter heads, filters.py:28-44)."""
  M, K, N = 16, 128, 30
  big = _rand(M, 4 * K, seed=1).float().cuda()
  W = _rand(K, N, seed=2).float().cuda()
  x = big[:, K:2 * K]
  ref = x.double().cpu() @ W.double().cpu()
  ybig = torch.zeros(M, 3 * N, device="cuda")
  nn.fc_fwd(x, W, None, mode=nn.FC_NOBIAS, out=ybig[:, N:2 * N])
  _close(ybig[:, N:2 * N], ref)
  assert float(ybig[:, :N].abs().max()) == 0 and float(ybig[:, 2 * N:].abs().max()) == 0
  dy = ybig[:, N:2 * N]
  _close(nn.fc_wgrad(x, dy), x.double().cpu().T @ ref)
  dxbig = torch.zeros(M, 4 * K, device="cuda")
  nn.fc_dgrad(dy, W, mul_act=x, out=dxbig[:, K:2 * K])
  _close(dxbig[:, K:2 * K], (ref @ W.double().cpu().T) * torch.where(x.double().cpu() > 0, 1.0, 0.2))


class _Cfg:
  exploration = 0.05; test_steps = 5; exploration_penalty = 0.05; filter_usage_penalty = 1.0
  all_reward = 1.0; critic_logit_multiplier = 0.05; discount_factor = 1.0; parameter_lr_mul = 1
  maximum_trajectory_length = 7; use_penalty = True


def test_stats_fwd_bwd_jvp(nn):
  from oracle import filters as F
  B, H, W = 5, 64, 64
  x = F.synth_images(B, H, W, seed=3).double()
  x[0, :8, :8] = 0.25                               # grey patch: 3-way max/min ties
  x[1, :4, :4, 0] = x[1, :4, :4, 1]                 # 2-way ties
  xr = x.clone().requires_grad_(True)
  st = N.critic_stats(xr)
  sd = nn.stats_fwd(x.float().cuda())
  _close(sd, st.detach(), tol=1e-5)
  gs = _rand(B, 3, seed=4)
  gdir = _rand(B, H, W, 3, seed=5)
  (gx,) = torch.autograd.grad(st, [xr], grad_outputs=gs, retain_graph=True)
  gd = nn.stats_bwd(x.float().cuda(), sd, gs.float().cuda(), g_direct=gdir.float().cuda())
  _close(gd, gx + gdir, tol=1e-5)
  u = _rand(B, H, W, 3, seed=6)
  jv = nn.stats_jvp(x.float().cuda(), sd, u.float().cuda())
  # <gs, J u> == <J^T gs, u>
  ref = torch.stack([(torch.autograd.grad(st[:, j].sum(), [xr], retain_graph=True)[0] * u).sum(dim=(1, 2, 3)) for j in range(3)], dim=1)
  _close(jv, ref, tol=1e-5)


@pytest.mark.parametrize("is_train", [1, 0])
def test_policy_head(nn, is_train):
  B, n = 64, 8
  cfg = _Cfg()
  logits = _rand(B, n, seed=1, scale=2.0).requires_grad_(True)
  u = torch.rand(B, 1, generator=torch.Generator().manual_seed(2), dtype=torch.float64)
  u[0] = 0.0                                        # pdf_sample quirk: id = -1
  states = torch.zeros(B, 3 + n, dtype=torch.float64)
  states[:, 2] = torch.arange(B) % 6
  states[:, 3:] = (torch.rand(B, n, generator=torch.Generator().manual_seed(3)) < 0.3).double()
  progress = 0.25
  pdf, ids, sur, ent, pen, ns = N.policy_head(logits, u, states, is_train, progress, cfg)
  d = nn.policy_head_fwd(logits.detach().float().cuda(), u[:, 0].float().cuda().contiguous(), states.float().cuda(),
                         is_train, progress, cfg)
  _close(d[0], pdf.detach(), 1e-5)
  assert torch.equal(d[1].cpu(), ids.to(torch.int32))
  if is_train:
    assert int(ids[0]) == -1
  _close(d[2], sur.detach()[:, 0], 1e-5); _close(d[3], ent.detach()[:, 0], 1e-5)
  _close(d[4], pen.detach()[:, 0], 1e-5); _close(d[5], ns.detach(), 1e-6)
  gsur, gpen = _rand(B, 1, seed=4), _rand(B, 1, seed=5)
  (gl,) = torch.autograd.grad([sur, pen], [logits], grad_outputs=[gsur, gpen])
  gd = nn.policy_head_bwd(logits.detach().float().cuda(), d[1], gsur[:, 0].float().cuda().contiguous(),
                          gpen[:, 0].float().cuda().contiguous(), progress, cfg)
  _close(gd, gl, 2e-5)


def test_overexposure_rl_losses_gp_adam(nn):
  from oracle import filters as F
  B = 32
  cfg = _Cfg()
  x = (F.synth_images(B, 64, 64, seed=9) * 6).double().requires_grad_(True)
  pen = (torch.clamp(x - 1, min=0) ** 2).mean(dim=(1, 2, 3))
  pd = nn.overexposure_fwd(x.detach().float().cuda())
  _close(pd, pen.detach(), 1e-5)
  gp = _rand(B, seed=1)
  (gx,) = torch.autograd.grad(pen, [x], grad_outputs=gp)
  gin = _rand(B, 64, 64, 3, seed=2)
  _close(nn.overexposure_bwd(x.detach().float().cuda(), gp.float().cuda(), g_in=gin.float().cuda()), gx + gin, 1e-5)
  # rl losses + seeds
  v = [_rand(B, 1, seed=10 + i).requires_grad_(True) for i in range(6)]
  ns = torch.zeros(B, 11, dtype=torch.float64)
  ns[:, 1] = (torch.arange(B) % 3 == 0).double()
  ns[:, 2] = torch.arange(B) % 10
  gl, vl = N.rl_losses(v[0], v[1], v[2], v[3], v[4], v[5], ns, cfg)
  seeds, losses = nn.rl_losses(*[t.detach().float().cuda().reshape(-1).contiguous() for t in v], ns.float().cuda(), cfg)
  _close(losses, torch.stack([gl, vl]).detach(), 1e-5)
  g_fl, g_nv, g_pen, g_sur = torch.autograd.grad(gl, [v[0], v[3], v[4], v[5]], retain_graph=True)
  (g_ov,) = torch.autograd.grad(vl, [v[2]])
  ref = torch.stack([g_fl, g_nv, g_ov, g_pen, g_sur])[:, :, 0]
  _close(seeds, ref, 1e-5)
  # gradient-penalty scaling
  g = _rand(B, 64, 64, 3, seed=20, scale=0.02).requires_grad_(True)
  norm = torch.sqrt(1e-6 + (g ** 2).sum(dim=(1, 2, 3)))
  gpv = 10.0 * torch.mean(torch.clamp(norm - 1, min=0) ** 2)
  (uref,) = torch.autograd.grad(gpv, [g])
  ud, nd = nn.gp_scale(g.detach().float().cuda(), 10.0)
  _close(nd, norm.detach(), 1e-5); _close(ud, uref, 2e-5)
  a = torch.rand(B, generator=torch.Generator().manual_seed(1))
  r, f = torch.rand(B, 8, 8, 3), torch.rand(B, 8, 8, 3)
  _close(nn.interpolate(r.cuda(), f.cuda(), a.cuda()), (r + a[:, None, None, None] * (f - r)).double(), 1e-6)
  # Adam vs the TF update rule
  n = 10007
  p, gr = _rand(n, seed=30), _rand(n, seed=31)
  m, vv = _rand(n, seed=32, scale=0.1), _rand(n, seed=33).abs() * 0.01
  lr, b1, b2, eps, t = 1.5e-5, 0.5, 0.9, 1e-8, 7
  lr_t = lr * (1 - b2 ** t) ** 0.5 / (1 - b1 ** t)
  m2 = b1 * m + (1 - b1) * gr; v2 = b2 * vv + (1 - b2) * gr * gr
  p2 = p - lr_t * m2 / (v2.sqrt() + eps)
  pc, mc, vc = p.float().cuda(), m.float().cuda(), vv.float().cuda()
  nn.adam(pc, gr.float().cuda(), mc, vc, torch.tensor([lr_t], device="cuda"), b1, b2, eps)
  _close(pc, p2, 1e-6); _close(mc, m2, 1e-6); _close(vc, v2, 1e-6)
