"""GPU parity: the CUDA filter step (through the C ABI) against the CPU oracle.

Tolerances (stated once, used everywhere below):
  forward pixels   |y - ref32| <= 1e-5 * max(|ref32|, 1e-4)        (north_star: <=1e-5 rel)
     ContrastFilter additionally gets the propagated effect of a 2-ulp difference in
     cosf(pi*l) between host and device libm, because the reference formula
     -cos(pi l)*0.5+0.5 cancels catastrophically for dark pixels; the CUDA
     result must also be no farther from the fp64 oracle than 4x the fp32 oracle is.
  image gradient   |gx - ref64| <= 1e-4 * max(|ref64|, 1e-3 * max|ref64|)
  param gradient   |gp - ref64| <= 1e-4 * max(|ref64|, 1e-3 * sum|terms|)
ref32 / ref64 = oracle/filters.py in float32 / float64 (pinned to the reference's Python by
tests/test_reference_golden.py; the CUDA path is also compared with that fixture directly in
tests/test_reference_golden_gpu.py)."""
import numpy as np
import pytest
import torch

from oracle import filters as F

pytestmark = pytest.mark.gpu
ALL = list(range(10))      # 0..7 = cfg.filters; 8 LevelFilter, 9 VignetFilter
SHAPES = [(4, 64, 64), (2, 33, 31), (3, 1, 1), (1, 7, 5), (2, 128, 96), (5, 2, 2)]


@pytest.fixture(scope="module")
def ops(built_lib):
  assert torch.cuda.is_available(), "GPU tests need a CUDA device"
  from exposure_b200 import ops as o
  return o


def _fwd_tol(fid, x, p, ref32):
  tol = 1e-5 * ref32.abs().clamp_min(1e-4)
  if fid == F.CT:
    lum = F.rgb2lum(x).clamp(0, 1)
    tol = tol + p.abs()[:, :, None, None] * x.abs() / (lum + 1e-6) * (2 * 2.0 ** -24)
  if fid == F.SP:
    # the fp32 reference turns a hue back into three ramps (2 - |6H - 2| ...): 6H carries up to ~6e-7 of ABSOLUTE rounding
    # error, i.e. up to 6e-7 * V * p in y.  The kernel computes the ramps in closed form ((c - min) / range, exact to ~1 ulp),
    # so it may differ from the fp32 restatement by that much while being closer to the fp64 one (checked below / by the bwd)
    V = x.clamp(max=1.0).amax(dim=-1, keepdim=True).abs()
    tol = tol + 1e-6 * V * p.abs().reshape(-1, 1, 1, 1)
  return tol


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("fid", ALL)
def test_forward_matches_oracle(ops, fid, shape):
  B, H, W = shape
  x = F.synth_images(B, H, W, seed=11 + fid)
  lg = F.synth_logits(fid, B)
  p32 = F.regress(fid, lg)
  ref32 = F.process(fid, x, p32)
  ref64 = F.process(fid, x.double(), F.regress(fid, lg.double()))
  params = ops.filter_regress_fwd(lg.cuda(), fid)
  assert torch.allclose(params[:, :F.NUM_PARAMS[fid]].cpu(), p32, rtol=2e-6, atol=1e-7)
  y = ops.filter_fwd(x.cuda(), params, fid).cpu()
  err = (y - ref32).abs()
  tol = _fwd_tol(fid, x, p32, ref32)
  assert (err <= tol).all(), "fid %d: max err/tol %.3g" % (fid, float((err / tol).max()))
  e_cuda = (y.double() - ref64).abs()
  e_ref = (ref32.double() - ref64).abs()
  assert (e_cuda <= 4 * e_ref + tol.double()).all()


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("fid", ALL)
def test_backward_matches_oracle(ops, fid, shape):
  B, H, W = shape
  x = F.synth_images(B, H, W, seed=23 + fid)
  lg = F.synth_logits(fid, B)
  g = torch.Generator().manual_seed(5)
  gy = torch.randn(B, H, W, 3, generator=g)
  p32 = F.regress(fid, lg)
  gx64, gp64 = F.process_bwd_analytic(fid, x.double(), p32.double(), gy.double())
  params = torch.zeros(B, 24)
  params[:, :p32.shape[1]] = p32
  gx, gp = ops.filter_bwd(x.cuda(), gy.cuda(), params.cuda(), fid)
  gx, gp = gx.cpu().double(), gp.cpu().double()[:, :F.NUM_PARAMS[fid]]
  tolx = 1e-4 * gx64.abs().clamp_min(1e-3 * float(gx64.abs().max()) + 1e-30)
  assert ((gx - gx64).abs() <= tolx).all(), "gx fid %d: %.3g" % (fid, float(((gx - gx64).abs() / tolx).max()))
  # scale of the summed terms (a cancelling sum cannot be more accurate than its terms)
  _, gp_abs = F.process_bwd_analytic(fid, x.double(), p32.double(), gy.double().abs())
  scale = gp_abs.abs().max(dim=1, keepdim=True).values
  tolp = 1e-4 * torch.maximum(gp64.abs(), 1e-3 * scale + 1e-30)
  assert ((gp - gp64).abs() <= tolp).all(), "gp fid %d: %.3g" % (fid, float(((gp - gp64).abs() / tolp).max()))
  # param-gradient-only backward (the reference's actual training need) gives the same gparams
  # (separately compiled instantiation: FMA contraction may differ in the last bit)
  _, gp2 = ops.filter_bwd(x.cuda(), gy.cuda(), params.cuda(), fid, need_gx=False)
  assert ((gp2.cpu().double()[:, :F.NUM_PARAMS[fid]] - gp).abs() <= 0.1 * tolp).all()


@pytest.mark.parametrize("fid", ALL)
def test_regressor_backward(ops, fid):
  B = 9
  lg = F.synth_logits(fid, B, seed=99)
  gp = torch.randn(B, 24)
  gp[:, F.NUM_PARAMS[fid]:] = 0
  ref = F.regress_bwd(fid, lg.double(), gp[:, :F.NUM_PARAMS[fid]].double())
  lgp = torch.zeros(B, 24)
  lgp[:, :F.NUM_PARAMS[fid]] = lg
  out = ops.filter_regress_bwd(lgp.cuda(), gp.cuda(), fid).cpu().double()
  assert torch.allclose(out[:, :F.NUM_PARAMS[fid]], ref, rtol=1e-4, atol=1e-7)
  assert (out[:, F.NUM_PARAMS[fid]:] == 0).all()


def test_per_image_filter_ids(ops):
  """agent.py:113-125: each image gets its own selected filter (ids tensor)."""
  B, H, W = 16, 32, 32
  x = F.synth_images(B, H, W, seed=3)
  ids = torch.arange(B, dtype=torch.int32) % 8
  lg = torch.randn(B, 24, generator=torch.Generator().manual_seed(1))
  params = ops.filter_regress_fwd(lg.cuda(), ids.cuda())
  y = ops.filter_fwd(x.cuda(), params, ids.cuda()).cpu()
  gy = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(2))
  gx, gp = ops.filter_bwd(x.cuda(), gy.cuda(), params, ids.cuda())
  for b in range(B):
    fid = int(ids[b])
    n = F.NUM_PARAMS[fid]
    # uniform-id launch on the single image must agree bit for bit with the select kernel
    pu = ops.filter_regress_fwd(lg[b:b + 1].cuda(), fid)
    yu = ops.filter_fwd(x[b:b + 1].cuda(), pu, fid).cpu()
    assert torch.equal(yu[0], y[b]), (b, fid)
    gxu, gpu = ops.filter_bwd(x[b:b + 1].cuda(), gy[b:b + 1].cuda(), pu, fid)
    assert torch.equal(gxu.cpu()[0], gx.cpu()[b]) and torch.equal(gpu.cpu()[0, :n], gp.cpu()[b, :n])
    ref = F.process(fid, x[b:b + 1], F.regress(fid, lg[b:b + 1, :n]))
    assert torch.allclose(y[b], ref[0], rtol=1e-5, atol=1e-6 if fid == F.SP else 1e-7) or fid == F.CT   # S+: see _fwd_tol


def test_scalar_and_vector_variants_agree(ops):
  from exposure_b200 import _cabi
  B, H, W = 3, 40, 36
  x = F.synth_images(B, H, W, seed=8).cuda()
  gy = torch.randn(B, H, W, 3, device="cuda")
  for fid in ALL:
    params = ops.filter_regress_fwd(F.synth_logits(fid, B).cuda(), fid)
    yv = ops.filter_fwd(x, params, fid, variant=_cabi.VARIANT_DIRECT)
    ys = ops.filter_fwd(x, params, fid, variant=_cabi.VARIANT_SCALAR)
    assert torch.equal(yv, ys), fid
    gxv, gpv = ops.filter_bwd(x, gy, params, fid, variant=_cabi.VARIANT_DIRECT)
    gxs, gps = ops.filter_bwd(x, gy, params, fid, variant=_cabi.VARIANT_SCALAR)
    assert torch.equal(gxv, gxs), fid
    assert torch.allclose(gpv, gps, rtol=1e-5, atol=1e-6), fid


@pytest.mark.parametrize("shape", [(3, 40, 36), (2, 64, 64), (1, 2, 2), (9, 512, 512), (5, 333, 500)])
def test_tma_variant_matches_direct(ops, shape):
  """EXP_VARIANT_TMA (persistent cp.async.bulk ring) == EXP_VARIANT_DIRECT: identical per-pixel
  code, so pixels and image gradients are bit-equal; parameter gradients differ only in the
  (deterministic) summation order.  Shapes cover partial tiles, 1 tile and many tiles per CTA."""
  from exposure_b200 import _cabi
  B, H, W = shape
  g = torch.Generator(device="cuda").manual_seed(3)
  x = torch.exp(torch.randn(B, H, W, 3, device="cuda", generator=g) - 3.2).clamp_(0, 4)
  x = torch.where(torch.rand(B, H, W, 3, device="cuda", generator=g) < 0.02, x * 8, x)
  gy = torch.randn(B, H, W, 3, device="cuda", generator=g)
  for fid in ALL:
    params = ops.filter_regress_fwd(F.synth_logits(fid, B).cuda(), fid)
    yd = ops.filter_fwd(x, params, fid, variant=_cabi.VARIANT_DIRECT)
    yt = ops.filter_fwd(x, params, fid, variant=_cabi.VARIANT_TMA)
    assert torch.equal(yd, yt), fid
    gxd, gpd = ops.filter_bwd(x, gy, params, fid, variant=_cabi.VARIANT_DIRECT)
    gxt, gpt = ops.filter_bwd(x, gy, params, fid, variant=_cabi.VARIANT_TMA)
    assert torch.equal(gxd, gxt), fid
    scale = gpd.abs().max() + 1e-6
    assert ((gpd - gpt).abs() <= 2e-5 * scale + 1e-4 * gpd.abs()).all(), (fid, (gpd - gpt).abs().max(), scale)
    _, gpt2 = ops.filter_bwd(x, gy, params, fid, need_gx=False, variant=_cabi.VARIANT_TMA)
    assert ((gpt - gpt2).abs() <= 2e-5 * scale + 1e-4 * gpt.abs()).all(), fid
    gxt3, gpt3 = ops.filter_bwd(x, gy, params, fid, variant=_cabi.VARIANT_TMA)
    assert torch.equal(gxt3, gxt) and torch.equal(gpt3, gpt), fid        # deterministic


def test_in_place_and_determinism(ops):
  B, H, W = 2, 64, 64
  x = F.synth_images(B, H, W, seed=4).cuda()
  gy = torch.randn(B, H, W, 3, device="cuda")
  for fid in ALL:
    params = ops.filter_regress_fwd(F.synth_logits(fid, B).cuda(), fid)
    y = ops.filter_fwd(x, params, fid)
    xi = x.clone()
    ops.filter_fwd(xi, params, fid, out=xi)
    assert torch.equal(xi, y)
    gx1, gp1 = ops.filter_bwd(x, gy, params, fid)
    gx2, gp2 = ops.filter_bwd(x, gy, params, fid)
    assert torch.equal(gx1, gx2) and torch.equal(gp1, gp2)      # deterministic reduction
    gi = gy.clone()
    ops.filter_bwd(x, gi, params, fid, gx_out=gi)
    assert torch.equal(gi, gx1)


def test_chain_fwd_bwd_matches_oracle(ops):
  """BASELINE configs[1] at oracle size: E,G,W,S+,T,Ct,BW,C applied in sequence, fwd+bwd."""
  ids = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]
  B, H, W = 4, 64, 64
  x = F.synth_images(B, H, W, seed=77)
  lgs = [F.synth_logits(f, B, seed=1000) * 0.5 for f in ids]
  gout = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(9))
  y64, gx64, glg64 = F.chain_fwd_bwd(ids, x.double(), [l.double() for l in lgs], gout.double())
  y32 = F.chain_fwd(ids, x, lgs)[-1]
  from exposure_b200.chain import FilterChain
  chain = FilterChain(ids)
  lg_dev = [l.cuda() for l in lgs]
  y = chain.forward(x.cuda(), lg_dev)
  gx, glg = chain.backward(gout.cuda())
  y, gx = y.cpu(), gx.cpu().double()
  # errors compound over 8 steps (pow/cos amplification): 5e-5 on the final pixels
  assert ((y - y32).abs() <= 5e-5 * y32.abs().clamp_min(1e-3)).all(), float(((y - y32).abs() / y32.abs().clamp_min(1e-3)).max())
  assert ((gx - gx64).abs() <= 1e-3 * gx64.abs().clamp_min(1e-3 * float(gx64.abs().max()))).all()
  for k, (a, b) in enumerate(zip(glg, glg64)):
    a = a.cpu().double()[:, :b.shape[1]]
    assert torch.allclose(a, b, rtol=2e-3, atol=2e-3 * float(b.abs().max()) + 1e-12), (k, (a - b).abs().max(), b.abs().max())


def test_full_size_properties(ops):
  """BASELINE configs[1] full size (64x512x512x3): size-independent properties."""
  B, H, W = 64, 512, 512
  g = torch.Generator(device="cuda").manual_seed(1)
  x = torch.exp(torch.randn(B, H, W, 3, device="cuda", generator=g) - 3.2).clamp_(0, 4)
  z = lambda fid: ops.filter_regress_fwd(torch.zeros(B, 24, device="cuda"), fid)
  # identities at default parameters
  assert torch.equal(ops.filter_fwd(x, z(F.E), F.E), x)
  assert torch.equal(ops.filter_fwd(x, z(F.CT), F.CT), x)
  assert torch.allclose(ops.filter_fwd(x, z(F.T), F.T), x.clamp(0, 1), rtol=0, atol=2e-7)
  assert torch.allclose(ops.filter_fwd(x, z(F.C), F.C), x.clamp(0, 1), rtol=0, atol=2e-7)
  # exact homogeneity of the linear filters under power-of-two scaling
  for fid in (F.E, F.W, F.BW):
    p = ops.filter_regress_fwd(torch.randn(B, 24, device="cuda", generator=g), fid)
    assert torch.equal(ops.filter_fwd(x * 0.5, p, fid), ops.filter_fwd(x, p, fid) * 0.5)
  # exposure round trip E(+p) then E(-p)
  lg = torch.randn(B, 24, device="cuda", generator=g)
  p = ops.filter_regress_fwd(lg, F.E)
  back = ops.filter_fwd(ops.filter_fwd(x, p, F.E), -p, F.E)
  assert torch.allclose(back, x, rtol=1e-6, atol=0)
  # parameter-gradient reduction: whole image == sum of its two halves; deterministic
  gy = torch.randn(B, H, W, 3, device="cuda", generator=g)
  for fid in (F.E, F.T, F.C):
    p = ops.filter_regress_fwd(torch.randn(B, 24, device="cuda", generator=g), fid)
    _, gp = ops.filter_bwd(x, gy, p, fid, need_gx=False)
    _, gpa = ops.filter_bwd(x[:, :256].contiguous(), gy[:, :256].contiguous(), p, fid, need_gx=False)
    _, gpb = ops.filter_bwd(x[:, 256:].contiguous(), gy[:, 256:].contiguous(), p, fid, need_gx=False)
    if fid == F.E:
      assert torch.allclose(gp, gpa + gpb, rtol=1e-4, atol=1e-3)
    _, gp_again = ops.filter_bwd(x, gy, p, fid, need_gx=False)
    assert torch.equal(gp, gp_again)


def test_fused_regressor_matches_separate_kernels(ops):
  """EXP_OPT_LOGITS: filter_param_regressor fused into the filter step == regress kernel + filter kernel."""
  from exposure_b200 import _cabi
  B, H, W = 3, 48, 40
  x = F.synth_images(B, H, W, seed=14).cuda()
  gy = torch.randn(B, H, W, 3, device="cuda")
  for fid in ALL:
    lg = F.synth_logits(fid, B).cuda()
    params = ops.filter_regress_fwd(lg, fid)
    for variant in (_cabi.VARIANT_DIRECT, _cabi.VARIANT_TMA, _cabi.VARIANT_SCALAR):
      y0 = ops.filter_fwd(x, params, fid, variant=variant)
      y1 = ops.filter_fwd(x, lg, fid, variant=variant, logits=True)
      assert torch.equal(y0, y1), (fid, variant)
      gx0, gp0 = ops.filter_bwd(x, gy, params, fid, variant=variant)
      gl0 = ops.filter_regress_bwd(lg, gp0, fid)
      gx1, gl1 = ops.filter_bwd(x, gy, lg, fid, variant=variant, logits=True)
      assert torch.equal(gx0, gx1), (fid, variant)
      n = F.NUM_PARAMS[fid]
      assert gl1.shape == (B, n) and torch.allclose(gl0, gl1, rtol=1e-6, atol=1e-7), (fid, variant)


def test_chain_graph_replay(ops):
  from exposure_b200.chain import FilterChain
  ids = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]
  B, H, W = 4, 64, 64
  x = F.synth_images(B, H, W, seed=3).cuda()
  lgs = [(F.synth_logits(f, B, seed=5) * 0.5).cuda() for f in ids]
  gout = torch.randn(B, H, W, 3, device="cuda")
  eager = FilterChain(ids)
  y = eager.forward(x, lgs).clone()
  gx, gl = eager.backward(gout)
  gx = gx.clone(); gl = [g.clone() for g in gl]
  ch = FilterChain(ids)
  ch.input_buffer(x.shape, x.device).copy_(x)
  ch.capture(lgs, gout)
  assert ch.graph_launches == 2 * len(ids)
  for _ in range(2):
    y2, gx2, gl2 = ch.replay()
  assert torch.equal(y2, y) and torch.equal(gx2, gx)
  for a, b in zip(gl2, gl):
    assert torch.equal(a, b)


def test_host_pipelined_chain_matches_resident(ops):
  from exposure_b200.chain import FilterChain, HostPipelinedChain
  ids = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]
  B, H, W = 8, 64, 64
  x = F.synth_images(B, H, W, seed=31)
  lgs = [(F.synth_logits(f, B, seed=7) * 0.5).cuda() for f in ids]
  gout = torch.randn(B, H, W, 3, device="cuda")
  ref = FilterChain(ids)
  y = ref.forward(x.cuda(), lgs).clone()
  _, gl = ref.backward(gout)
  hx = x.pin_memory()
  hy = torch.empty(B, H, W, 3).pin_memory()
  hg = [torch.empty(B, F.NUM_PARAMS[f]).pin_memory() for f in ids]
  pipe = HostPipelinedChain(ids, B, H, W, torch.device("cuda"), chunks=4)
  for _ in range(3):                                   # buffers are reused across steps
    hy.zero_()
    pipe.step(hx, lgs, gout, hy, hg)
    assert torch.equal(hy, y.cpu())
    for a, b in zip(hg, gl):
      assert torch.equal(a, b.cpu())
  # enqueue-only steps (consecutive steps overlap), results after wait()
  hy.zero_()
  for _ in range(3):
    pipe.step(hx, lgs, gout, hy, hg, wait=False)
  pipe.wait()
  assert torch.equal(hy, y.cpu())
  for a, b in zip(hg, gl):
    assert torch.equal(a, b.cpu())


MASK_KW = dict(maximum_sharpness=1.5, minimum_strength=0.3)


@pytest.mark.parametrize("masking", [True, False])
@pytest.mark.parametrize("shape", [(2, 64, 64), (2, 24, 40), (3, 21, 9), (1, 130, 70)])
@pytest.mark.parametrize("fid", ALL)
def test_masked_apply_matches_oracle(ops, fid, shape, masking):
  """Filter.apply with cfg.masking (filters.py:62-99, 110-148; Vignet 354-396): pixels, mask,
  image / filter-logit / mask-logit gradients against the oracle (autograd of the restatement)."""
  B, H, W = shape
  x = F.synth_images(B, H, W, seed=60 + fid)
  lg = F.synth_logits(fid, B) * 0.7
  nm = 5 if fid == F.VG else 6
  ml = torch.randn(B, nm, generator=torch.Generator().manual_seed(3 + fid)) * 0.8
  gy = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(9 + fid))
  ref32 = F.apply_masked(fid, x, lg, ml, masking, **MASK_KW)
  mask32 = F.get_mask(fid, x, ml, masking, **MASK_KW).expand(B, H, W, 1)
  gx64, gl64, gm64 = F.apply_masked_bwd_autograd(fid, x.double(), lg.double(), ml.double(), gy.double(), masking, **MASK_KW)
  lgp = torch.zeros(B, 24); lgp[:, :lg.shape[1]] = lg
  mlp = torch.zeros(B, 6); mlp[:, :nm] = ml
  xd, lgd, mld, gyd = x.cuda(), lgp.cuda(), mlp.cuda(), gy.cuda()
  y, mask = ops.filter_masked_fwd(xd, lgd, mld, fid, 1.5, 0.3, masking, want_mask=True, logits=True)
  assert torch.allclose(mask.cpu(), mask32, rtol=0, atol=3e-6)
  assert torch.equal(ops.filter_mask(xd, mld, fid, 1.5, 0.3, masking), mask)
  proc = F.process(fid, x, F.regress(fid, lg))
  tol = 1e-5 * ref32.abs().clamp_min(1e-4) + 4e-6 * (proc - x).abs()
  if fid in (F.CT, F.SP):                            # the filter's own extra slack (see _fwd_tol)
    tol = tol + _fwd_tol(fid, x, F.regress(fid, lg), proc) - 1e-5 * proc.abs().clamp_min(1e-4)
  err = (y.cpu() - ref32).abs()
  assert (err <= tol).all(), float((err / tol).max())
  gx, gl, gm = ops.filter_masked_bwd(xd, gyd, lgd, mld, fid, 1.5, 0.3, masking, logits=True)
  if fid != F.SP:      # TF defines no S+ image gradient (closed form checked in test_backward_matches_oracle)
    e = (gx.cpu().double() - gx64).abs()
    assert (e <= 1e-4 * gx64.abs().clamp_min(1e-3 * float(gx64.abs().max()) + 1e-30)).all(), float(e.max())
  n = F.NUM_PARAMS[fid]
  for got, ref in ((gl.cpu()[:, :n], gl64), (gm.cpu()[:, :nm], gm64)):
    eg = (got.double() - ref).abs()
    scale = ref.abs().max(dim=1, keepdim=True).values.clamp_min(1e-6)
    assert (eg <= 2e-4 * scale).all(), (fid, float((eg / scale).max()))
  # parameter-only backward and determinism
  _, gl2, gm2 = ops.filter_masked_bwd(xd, gyd, lgd, mld, fid, 1.5, 0.3, masking, need_gx=False, logits=True)
  assert torch.allclose(gl2, gl, rtol=1e-5, atol=1e-6) and torch.allclose(gm2, gm, rtol=1e-5, atol=1e-6)
  gx3, gl3, gm3 = ops.filter_masked_bwd(xd, gyd, lgd, mld, fid, 1.5, 0.3, masking, logits=True)
  assert torch.equal(gx3, gx) and torch.equal(gl3, gl) and torch.equal(gm3, gm)
  if not masking and fid != F.VG:      # mask == 1: the masked kernels reduce to the plain step
    assert torch.allclose(y, ops.filter_fwd(xd, lgd, fid, logits=True), rtol=1e-6, atol=1e-7)
    assert float(gm.abs().max()) == 0.0


def test_masked_per_image_ids(ops):
  """Per-image filter ids through the masked kernels == uniform launches image by image."""
  B, H, W = 10, 32, 32
  x = F.synth_images(B, H, W, seed=5).cuda()
  ids = torch.arange(B, dtype=torch.int32).cuda()
  lg = (torch.randn(B, 24, generator=torch.Generator().manual_seed(1)) * 0.7).cuda()
  ml = (torch.randn(B, 6, generator=torch.Generator().manual_seed(2)) * 0.8).cuda()
  gy = torch.randn(B, H, W, 3, device="cuda")
  y = ops.filter_masked_fwd(x, lg, ml, ids, 1.0, 0.3, True, logits=True)
  gx, gl, gm = ops.filter_masked_bwd(x, gy, lg, ml, ids, 1.0, 0.3, True, logits=True)
  for b in range(B):
    s = slice(b, b + 1)
    yu = ops.filter_masked_fwd(x[s].contiguous(), lg[s].contiguous(), ml[s].contiguous(), b, 1.0, 0.3, True, logits=True)
    assert torch.equal(yu[0], y[b]), b
    gxu, glu, gmu = ops.filter_masked_bwd(x[s].contiguous(), gy[s].contiguous(), lg[s].contiguous(), ml[s].contiguous(),
                                          b, 1.0, 0.3, True, logits=True)
    assert torch.equal(gxu[0], gx[b]) and torch.equal(glu[0], gl[b]) and torch.equal(gmu[0], gm[b]), b


def test_errors_are_loud(ops):
  from exposure_b200._cabi import ExposureLibError
  x = torch.zeros(1, 3, 3, 3, device="cuda")
  p = torch.zeros(1, 24, device="cuda")
  with pytest.raises(ExposureLibError):
    ops.filter_fwd(x, p, 11)
  with pytest.raises(ExposureLibError):
    ops.filter_fwd(x, p, 0, variant=1)       # DIRECT needs H*W % 4 == 0
  with pytest.raises(ValueError):
    ops.filter_fwd(x.cpu(), p, 0)            # no CPU fallback


def test_satplus_gradient_on_exact_channel_ties(ops):
  """at pixels where two channels tie for the max / min, and at grey pixels, the CUDA backward
  follows the documented rule (first channel in R, G, B order is THE max / min; grey = TF's hue-0 branch) -- the
  oracle's analytic gradient, which tests/test_oracle_filters.py pins to one-sided derivatives of the
  reference-pinned forward."""
  px = []
  for hi, lo in ((0.8, 0.3), (0.4, 0.1), (0.9, 0.6), (0.3, 0.05)):
    px += [(hi, hi, lo), (hi, lo, hi), (lo, hi, hi), (hi, lo, lo), (lo, hi, lo), (lo, lo, hi), (hi, hi, hi)]
  x = torch.tensor(px, dtype=torch.float32).reshape(1, 4, 7, 3)
  lg = torch.tensor([[0.8]])
  p32 = F.regress(F.SP, lg)
  g = torch.Generator().manual_seed(3)
  gy = torch.randn(x.shape, generator=g)
  gx64, gp64 = F.process_bwd_analytic(F.SP, x.double(), p32.double(), gy.double())
  params = torch.zeros(1, 24)
  params[:, :1] = p32
  gx, gp = ops.filter_bwd(x.cuda(), gy.cuda(), params.cuda(), F.SP)
  assert ((gx.cpu().double() - gx64).abs() <= 1e-4 * gx64.abs().clamp_min(1e-3 * float(gx64.abs().max()))).all()
  assert abs(float(gp[0, 0]) - float(gp64[0, 0])) <= 1e-4 * max(abs(float(gp64[0, 0])), 1e-3)
