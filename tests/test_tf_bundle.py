"""CPU: the TF-free bundle reader against the reference's shipped pretrained checkpoint (only
available in the build container; skipped on the GPU box where /root/reference does not exist),
and the oracle policy driven by those weights reproduces the survey's plausibility trajectory
(dark RAW -> Exposure first)."""
import os

import numpy as np
import pytest
import torch

PREFIX = "/root/reference/models/example/pretrained/model.ckpt-20000"
pytestmark = pytest.mark.skipif(not os.path.exists(PREFIX + ".index"), reason="reference checkpoint not present")


def test_index_entries_and_known_scalars():
  from exposure_b200.tf_bundle import load_bundle, read_index
  idx = read_index(PREFIX)
  assert idx["generator/Conv/weights"]["shape"] == (4, 4, 14, 32)
  assert idx["critic/Conv/weights"]["shape"] == (4, 4, 6, 32)
  assert idx["rl_value/critic/Conv/weights"]["shape"] == (4, 4, 17, 32)
  assert idx["generator/filter_7/fc2/weights"]["shape"] == (128, 30)
  assert idx["generator/action_selection/selector_fc2/weights"]["shape"] == (128, 8)
  w = load_bundle(PREFIX)
  n_params = sum(v.size for k, v in w.items() if v.dtype == np.float32 and v.ndim > 0)
  assert n_params == 8561762 and len(w) == 82     # 8 561 768 incl. the 6 scalar entries
  allv = load_bundle(PREFIX, include_optimizer_slots=True)
  counters = sorted(int(v) for k, v in allv.items() if k.startswith("Variable") and v.ndim == 0)
  assert counters == [20099, 20099, 104655]       # net.py:312-322 schedule for 20000 iterations
  ema = [float(v) for k, v in allv.items() if "ExponentialMovingAverage" in k and v.ndim == 0 and v.dtype == np.float32]
  assert any(abs(e - 13.55344) < 1e-4 for e in ema)


def test_pretrained_policy_brightens_dark_raw_first():
  """Appendix B of SURVEY: with the shipped weights (dropout replaced by its expectation) the
  oracle policy's first action on a dark linear image is Exposure with a positive EV."""
  from exposure_b200.tf_bundle import load_bundle
  from oracle import filters as OF
  from oracle import train_step as OT
  from exposure_b200.trainer import default_cfg
  w = {k: torch.from_numpy(np.array(v)).double() for k, v in load_bundle(PREFIX).items() if v.dtype == np.float32 and v.ndim > 0}
  cfg = default_cfg()
  img = OF.synth_images(2, 64, 64, seed=3, stress=False).double()          # median ~0.04: dark RAW statistics
  states = torch.zeros(2, 11, dtype=torch.float64)
  ones = torch.full((2, 4, 4, 256), 0.5, dtype=torch.float64)             # x * mask / keep with mask == keep
  out, new_states, sur, pen, ids, pdf = OT.agent_generator(w, img, states, torch.full((2,), 0.5, dtype=torch.float64),
                                                           ones * 2 * 0.5 * 2, ones * 2 * 0.5 * 2, 0, 0.0, cfg)
  assert ids.tolist() == [0, 0], (ids, pdf)                                # Exposure
  assert float(out.mean()) > 2 * float(img.mean())                         # ~ +1.4 .. +2 EV


def test_bundle_writer_checksums_match_the_shipped_checkpoint():
  """CRC-32C + TF's mask reproduce the checksums TF 1.6 stored in the reference's own checkpoint
  (golden bytes extracted by tests/golden/make_bundle_golden.py), and the header / footer layout the
  writer emits is the one found there."""
  import os
  import struct
  import numpy as np
  from exposure_b200 import tf_bundle as tb
  g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bundle_golden.npz"))
  assert tb.crc32c(b"123456789") == 0xE3069283                       # CRC-32C check value
  assert tb._mask_crc(tb.crc32c(g["block"].tobytes())) == int(g["block_crc"])
  assert tb._mask_crc(tb.crc32c(g["tensor"].tobytes())) == int(g["tensor_crc"])
  assert bytes(g["header"]) == tb._field(1, 0, tb._put_varint(1)) + tb._field(3, 2, tb._put_varint(2) + tb._field(1, 0, tb._put_varint(1)))
  assert struct.unpack_from("<Q", bytes(g["footer"]), 40)[0] == tb._MAGIC


def test_bundle_write_read_round_trip(tmp_path):
  import numpy as np
  from exposure_b200 import tf_bundle as tb
  rng = np.random.default_rng(0)
  t = {"generator/Conv/weights": rng.standard_normal((4, 4, 14, 32)).astype(np.float32),
       "generator/Conv/biases": np.zeros(32, np.float32), "Variable": np.int32(20099).reshape(()),
       "OptimizeLoss/beta1_power": np.float32(0.5 ** 7).reshape(())}
  for i in range(300):                                                # several table blocks
    t["pad/%03d/Adam_1" % i] = rng.standard_normal((i % 5 + 1, 3)).astype(np.float32)
  p = str(tmp_path / "model.ckpt-1")
  tb.save_bundle(p, t)
  assert tb.verify_bundle(p) == len(t)
  back = tb.load_bundle(p, include_optimizer_slots=True)
  assert set(back) == set(t)
  for k in t:
    assert back[k].shape == np.asarray(t[k]).shape and back[k].dtype == np.asarray(t[k]).dtype and np.array_equal(back[k], t[k])
