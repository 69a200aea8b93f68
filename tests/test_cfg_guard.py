"""CPU: configuration branches of the reference that the hot path does not implement are refused loudly."""
import pytest


def test_unsupported_config_branches_raise():
  from exposure_b200.trainer import Trainer, default_cfg
  Trainer._check_cfg(default_cfg())                         # the shipped configuration passes
  for key, value in (("gan", "ls"), ("supervised", True), ("use_TD", False), ("clamp", True),
                     ("shared_feature_extractor", False), ("img_include_states", False), ("gradient_penalty_lambda", 0)):
    cfg = default_cfg()
    cfg[key] = value
    with pytest.raises(NotImplementedError, match=key.split("_")[0]):
      Trainer._check_cfg(cfg)


def test_filter_list_and_network_sizes_are_checked():
  """a cfg with reordered / added filters or different network sizes must not silently train the
  shipped configuration (agent.py:43-59 derives action ids, fc2 widths and the state vector from cfg.filters)."""
  from exposure_b200 import filters as F
  from exposure_b200.trainer import Trainer, default_cfg
  cfg = default_cfg()
  cfg.filters = list(reversed(cfg.filters))
  with pytest.raises(NotImplementedError, match="cfg.filters"):
    Trainer._check_cfg(cfg)
  cfg = default_cfg()
  cfg.filters = cfg.filters + [F.LevelFilter]
  with pytest.raises(NotImplementedError, match="cfg.filters"):
    Trainer._check_cfg(cfg)
  for key, value in (("num_state_dim", 12), ("base_channels", 16), ("fc1_size", 64), ("feature_extractor_dims", 2048),
                     ("source_img_size", 32), ("curve_steps", 4)):
    cfg = default_cfg()
    cfg[key] = value
    with pytest.raises(NotImplementedError, match=key):
      Trainer._check_cfg(cfg)


def test_filter_ranges_cross_the_abi(built_lib):
  """cfg.exposure_range / gamma_range / tone_curve_range / color_curve_range are handed to the library
  (filters.py:179, 202, 261, 309) instead of being asserted equal to baked constants."""
  from exposure_b200 import _cabi
  from exposure_b200.trainer import default_cfg
  try:
    cfg = default_cfg()
    cfg.exposure_range = 2.0; cfg.gamma_range = 2.5; cfg.tone_curve_range = (0.25, 3); cfg.color_curve_range = (0.8, 1.3)
    _cabi.set_filter_ranges(cfg)
    r = _cabi.get_filter_ranges()
    assert abs(r["exposure_range"] - 2.0) < 1e-6 and abs(r["gamma_range"] - 2.5) < 1e-5
    assert (round(r["tone_lo"], 6), round(r["tone_hi"], 6)) == (0.25, 3.0)
    assert (round(r["color_lo"], 6), round(r["color_hi"], 6)) == (0.8, 1.3)
    cfg.curve_steps = 16                                   # the curve kernels are built for 8 knots: refused, not ignored
    with pytest.raises(_cabi.ExposureLibError, match="curve_steps"):
      _cabi.set_filter_ranges(cfg)
    cfg.curve_steps = 8; cfg.color_curve_range = (1.1, 1.3)     # tanh_range(initial=1) needs lo < 1 < hi (util.py:285)
    with pytest.raises(_cabi.ExposureLibError, match="initial value 1"):
      _cabi.set_filter_ranges(cfg)
  finally:
    _cabi.set_filter_ranges(None)
  r = _cabi.get_filter_ranges()
  assert abs(r["exposure_range"] - 3.5) < 1e-6 and abs(r["gamma_range"] - 3.0) < 1e-5
