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
