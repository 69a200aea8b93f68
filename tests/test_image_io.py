"""CPU: image decode + linearisation around the eval path (net.py:731-747) on the reference's
sample inputs (present only in the build container)."""
import os

import numpy as np
import pytest

SAMPLES = "/root/reference/models/sample_inputs"
pytestmark = pytest.mark.skipif(not os.path.isdir(SAMPLES), reason="reference sample inputs not present")


def test_tif_and_png_linearisation(tmp_path):
  from exposure_b200.evaluate import load_linear_image, save_png
  a = load_linear_image(os.path.join(SAMPLES, "A.tif"))
  assert a.dtype == np.float32 and a.shape == (333, 500, 3)
  assert 0.02 < float(a.mean()) < 0.1 and float(a.max()) <= 1.0          # dark linear RAW (SURVEY 8d statistics)
  b = load_linear_image(os.path.join(SAMPLES, "D-8bit-png.png"))
  assert abs(float(b.max()) - 0.5) < 1e-6                                 # "/ (2 * max)" mimics RAW exposure
  out = str(tmp_path / "x.png")
  save_png(out, np.clip(a * 4, 0, 1))
  assert os.path.getsize(out) > 1000


def test_thumbnail_pipeline_matches_reference_code():
  """load_linear_image + center_thumbnail against the thumbnails the reference's own util.py functions made
  (linearize_ProPhotoRGB, get_image_center, cv2.resize(..., (64, 64)) as in net.py:726,779; stored as
  `thumbs` in tests/golden/reference_golden.npz by make_reference_golden.py)."""
  import torch
  from exposure_b200.evaluate import center_thumbnail, load_linear_image
  gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_golden.npz"))
  for k, name in enumerate(("A.tif", "H.tif")):
    lin = load_linear_image(os.path.join(SAMPLES, name))
    thumb = center_thumbnail(torch.from_numpy(lin)[None], 64)[0].numpy()
    want = gold["thumbs"][k]
    assert thumb.shape == want.shape == (64, 64, 3)
    assert float(np.abs(thumb - want).max()) < 2e-6, name
