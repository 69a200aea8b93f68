"""The reference-facing Filter registry (exposure_b200/filters.py) reads like the reference's own
usage: classes from cfg.filters instantiated as x(net, cfg), apply() with specified parameters or
with image features, autograd through process()."""
import pytest
import torch

from oracle import filters as OF

pytestmark = pytest.mark.gpu


def test_registry_apply_and_autograd(built_lib):
  from exposure_b200 import filters as FL
  from exposure_b200.trainer import default_cfg
  cfg = default_cfg()
  cfg.filters = [FL.ExposureFilter, FL.GammaFilter, FL.ImprovedWhiteBalanceFilter, FL.SaturationPlusFilter,
                 FL.ToneFilter, FL.ContrastFilter, FL.WNBFilter, FL.ColorFilter]      # config_example.py:22-25
  B = 4
  net = OF.synth_images(B, 64, 64, seed=2).cuda()
  filters = [x(net, cfg) for x in cfg.filters]                                        # agent.py:45
  assert [f.get_short_name() for f in filters] == OF.FILTER_NAMES[:8]
  assert [f.get_num_filter_parameters() for f in filters] == OF.NUM_PARAMS[:8]
  for j, f in enumerate(filters):
    logits = OF.synth_logits(j, B).cuda().requires_grad_(True)
    param = f.filter_param_regressor(logits)
    ref_p = OF.regress(j, logits.detach().cpu())
    assert torch.allclose(param.detach().cpu().reshape(B, -1), ref_p, rtol=2e-6, atol=1e-7)
    low, high, dbg = f.apply(net, specified_parameter=param, high_res=net[:, :32].contiguous())
    ref = OF.process(j, net.cpu(), ref_p)
    tol = 1e-5 * ref.abs().clamp_min(1e-4) + (2e-4 * ref.abs() if j == OF.CT else 0)
    if j == OF.SP:      # closed-form ramps vs the fp32 hue round trip of the restatement (tests/test_filters_gpu.py _fwd_tol)
      tol = tol + 1e-6 * net.cpu().clamp(max=1.0).amax(dim=-1, keepdim=True).abs() * ref_p.abs().reshape(-1, 1, 1, 1)
    assert ((low.detach().cpu() - ref).abs() <= tol).all(), j
    assert high.shape == (B, 32, 64, 3) and "filter_parameters" in dbg and "mask" in dbg
    low.sum().backward()                                                              # tf.gradients through process + regressor
    gl = OF.regress_bwd(j, logits.detach().cpu().double(),
                        OF.process_bwd_analytic(j, net.cpu().double(), ref_p.double(), torch.ones(B, 64, 64, 3, dtype=torch.float64))[1])
    assert torch.allclose(logits.grad.cpu().double(), gl, rtol=2e-3, atol=2e-3 * float(gl.abs().max()) + 1e-9), j


def test_apply_with_features(built_lib):
  from exposure_b200 import filters as FL
  from exposure_b200.trainer import default_cfg
  cfg = default_cfg()
  B = 3
  net = OF.synth_images(B, 64, 64, seed=4).cuda()
  f = FL.ToneFilter(net, cfg)
  g = torch.Generator().manual_seed(0)
  v = {"fc1/weights": (torch.randn(4096, 128, generator=g) * 0.02).cuda(), "fc1/biases": torch.zeros(128).cuda(),
       "fc2/weights": (torch.randn(128, 14, generator=g) * 0.1).cuda(), "fc2/biases": torch.zeros(14).cuda()}
  f.bind_variables(v)
  feats = torch.randn(B, 4096, generator=g).cuda()
  low, _, dbg = f.apply(net, img_features=feats)
  h = feats.cpu() @ v["fc1/weights"].cpu()
  h = 0.6 * h + 0.4 * h.abs()
  o = h @ v["fc2/weights"].cpu()
  ref = OF.apply_filter(OF.T, net.cpu(), o[:, :8])
  assert torch.allclose(low.cpu(), ref, rtol=1e-4, atol=1e-6)


def test_masking_level_and_vignet_through_the_registry(built_lib):
  """cfg.masking = True (filters.py:62-148): apply() with image features runs the masked step kernel;
  LevelFilter / VignetFilter (filters.py:449-464, 341-396) instantiate like the shipped eight."""
  from exposure_b200 import filters as FL
  from exposure_b200.trainer import default_cfg
  cfg = default_cfg()
  cfg.masking = True
  cfg.maximum_sharpness, cfg.minimum_strength = 1, 0.3
  B = 3
  net = OF.synth_images(B, 64, 64, seed=6).cuda()
  g = torch.Generator().manual_seed(0)
  feats = torch.randn(B, 4096, generator=g).cuda()
  for cls, fid in ((FL.ContrastFilter, OF.CT), (FL.LevelFilter, OF.LE), (FL.VignetFilter, OF.VG)):
    f = cls(net, cfg)
    n, nm = f.get_num_filter_parameters(), f.get_num_mask_parameters()
    assert (n, nm) == (OF.NUM_PARAMS[fid], 5 if fid == OF.VG else 6)
    v = {"fc1/weights": (torch.randn(4096, 128, generator=g) * 0.02).cuda().requires_grad_(True),
         "fc1/biases": torch.zeros(128).cuda(),
         "fc2/weights": (torch.randn(128, n + nm, generator=g) * 0.1).cuda(), "fc2/biases": torch.zeros(n + nm).cuda()}
    f.bind_variables(v)
    low, high, dbg = f.apply(net, img_features=feats, high_res=net[:, :32].contiguous())
    h = feats.cpu() @ v["fc1/weights"].detach().cpu()
    h = 0.6 * h + 0.4 * h.abs()
    o = h @ v["fc2/weights"].cpu()
    ref = OF.apply_masked(fid, net.cpu(), o[:, :n], o[:, n:], True)
    assert torch.allclose(low.detach().cpu(), ref, rtol=2e-4, atol=2e-6), fid
    ref_mask = OF.get_mask(fid, net.cpu(), o[:, n:], True)
    assert torch.allclose(dbg["mask"].cpu(), ref_mask[0], rtol=0, atol=1e-5)
    assert high.shape == (B, 32, 64, 3) and f.high_res_mask.shape == (B, 32, 64, 1)
  # the `specified_parameter` branch asserts masking is off, like the reference (filters.py:72)
  with pytest.raises(AssertionError):
    FL.ExposureFilter(net, cfg).apply(net, specified_parameter=torch.zeros(B, 1).cuda())


def test_cfg_ranges_reach_the_kernels(built_lib, monkeypatch):
  """cfg.exposure_range / gamma_range / tone_curve_range / color_curve_range are read like the reference reads
  them (filters.py:179, 202, 261, 309), not asserted equal to baked constants: a cfg with other ranges goes
  through the registry and through the fused regressor of the step kernels (EXP_OPT_LOGITS)."""
  from exposure_b200 import _cabi, ops
  from exposure_b200 import filters as FL
  from exposure_b200.trainer import default_cfg
  cfg = default_cfg()
  cfg.exposure_range = 2.0; cfg.gamma_range = 2.5; cfg.tone_curve_range = (0.25, 3.0); cfg.color_curve_range = (0.8, 1.3)
  monkeypatch.setattr(OF, "EXPOSURE_RANGE", 2.0)
  monkeypatch.setattr(OF, "GAMMA_RANGE", 2.5)
  monkeypatch.setattr(OF, "TONE_CURVE_RANGE", (0.25, 3.0))
  monkeypatch.setattr(OF, "COLOR_CURVE_RANGE", (0.8, 1.3))
  B = 4
  net = OF.synth_images(B, 32, 32, seed=12).cuda()
  try:
    for cls, j in ((FL.ExposureFilter, OF.E), (FL.GammaFilter, OF.G), (FL.ToneFilter, OF.T), (FL.ColorFilter, OF.C)):
      f = cls(net, cfg)
      logits = OF.synth_logits(j, B).cuda().requires_grad_(True)
      param = f.filter_param_regressor(logits)
      ref_p = OF.regress(j, logits.detach().cpu())
      assert torch.allclose(param.detach().cpu().reshape(B, -1), ref_p, rtol=3e-6, atol=1e-7), j
      low, _, _ = f.apply(net, specified_parameter=param)
      ref = OF.process(j, net.cpu(), ref_p)
      assert ((low.detach().cpu() - ref).abs() <= 1e-5 * ref.abs().clamp_min(1e-4)).all(), j
      low.sum().backward()
      gl = OF.regress_bwd(j, logits.detach().cpu().double(),
                          OF.process_bwd_analytic(j, net.cpu().double(), ref_p.double(), torch.ones(B, 32, 32, 3, dtype=torch.float64))[1])
      assert torch.allclose(logits.grad.cpu().double(), gl, rtol=2e-3, atol=2e-3 * float(gl.abs().max()) + 1e-9), j
      # fused regressor inside the step kernels (the chain / train path)
      lg = torch.zeros(B, ops.PSTRIDE, device="cuda"); lg[:, :logits.shape[1]] = logits.detach()
      y = ops.filter_fwd(net, lg, j, logits=True)
      assert ((y.cpu() - ref).abs() <= 1e-5 * ref.abs().clamp_min(1e-4)).all(), j
      _, glk = ops.filter_bwd(net, torch.ones_like(net), lg, j, need_gx=False, logits=True)
      assert torch.allclose(glk[:, :logits.shape[1]].cpu().double(), gl, rtol=2e-3, atol=2e-3 * float(gl.abs().max()) + 1e-9), j
  finally:
    _cabi.set_filter_ranges(None)


def test_specified_parameter_of_batch_one_broadcasts(built_lib):
  """apply(specified_parameter=[1, n]) broadcasts over the batch like the reference's
  param[:, None, None, :] (filters.py:62-99); a mismatching batch raises instead of being reinterpreted."""
  from exposure_b200 import filters as FL
  from exposure_b200.trainer import default_cfg
  cfg = default_cfg()
  B = 4
  net = OF.synth_images(B, 16, 16, seed=3).cuda()
  f = FL.ToneFilter(net, cfg)
  p1 = f.filter_param_regressor(OF.synth_logits(OF.T, 1).cuda())          # [1,1,1,1,8]
  low, _, _ = f.apply(net, specified_parameter=p1)
  ref = OF.process(OF.T, net.cpu(), p1.reshape(1, 8).cpu().expand(B, 8))
  assert torch.allclose(low.cpu(), ref, rtol=1e-5, atol=1e-7)
  p2 = f.filter_param_regressor(OF.synth_logits(OF.T, 2).cuda())
  with pytest.raises(ValueError, match="batch"):
    f.apply(net, specified_parameter=p2)


def test_unselected_rows_are_black_without_prezeroing(built_lib):
  """pdf_sample yields id -1 when the uniform draw is exactly 0 (pdf_sample_layer.py:5-10: all-zero
  one-hot row -> the reference's one-hot sum is a black image).  The step kernels write those zeros themselves:
  poisoned (NaN) output buffers come back clean, forward and backward, plain and masked."""
  from exposure_b200 import ops
  B, H, W = 6, 16, 16
  x = OF.synth_images(B, H, W, seed=8).cuda()
  ids = torch.tensor([0, -1, 4, -1, 7, 3], dtype=torch.int32, device="cuda")
  params = torch.zeros(B, ops.PSTRIDE, device="cuda")
  for j in (0, 2, 4, 5):
    params[j] = ops.filter_regress_fwd(torch.zeros(1, ops.PSTRIDE, device="cuda"), int(ids[j]))[0]
  nan = lambda *s: torch.full(s, float("nan"), device="cuda")
  y = ops.filter_fwd(x, params, ids, out=nan(B, H, W, 3))
  assert torch.isfinite(y).all() and float(y[1].abs().max()) == 0.0 and float(y[3].abs().max()) == 0.0
  assert float(y[0].abs().max()) > 0
  gx, gp = ops.filter_bwd(x, torch.ones_like(x), params, ids, gx_out=nan(B, H, W, 3), gparams_out=nan(B, ops.PSTRIDE))
  assert torch.isfinite(gx).all() and float(gx[1].abs().max()) == 0.0 and float(gp[1].abs().max()) == 0.0
  assert torch.isfinite(gp[[1, 3]]).all()
  ym = ops.filter_masked_fwd(x, params, torch.zeros(B, 6, device="cuda"), ids, out=nan(B, H, W, 3))
  assert torch.isfinite(ym).all() and float(ym[3].abs().max()) == 0.0


def test_policy_forward_with_zero_noise_is_black_not_garbage(built_lib):
  """The same quirk through PolicyNet.forward (agent.py:113-125): noise == 0 -> id -1 -> out row == 0."""
  from exposure_b200.trainer import Trainer
  t = Trainer(device=torch.device("cuda", 0), seed=0)
  B = 4
  img = OF.synth_images(B, 64, 64, seed=9).cuda()
  states = torch.zeros(B, 11, device="cuda")
  noise, drop_f, drop_s, _ = t.draw(B, torch.Generator(device="cuda").manual_seed(1))
  noise[2] = 0.0
  # poison the caching allocator's free blocks so that an unwritten output would show
  junk = torch.full((B, 64, 64, 3), float("nan"), device="cuda"); del junk
  c = t.policy.forward(img, states, noise, drop_f, drop_s, 1, 0.1, t.cfg)
  assert int(c.ids[2]) == -1
  assert torch.isfinite(c.out).all() and float(c.out[2].abs().max()) == 0.0
