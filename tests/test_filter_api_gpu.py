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
