"""The host-side visual debugger (exposure_b200/visualize.py, Filter.visualize_* of the mirror classes and
the agent's `debugger` closure) against canvases drawn by the reference's OWN code
(tests/golden/reference_golden.npz section 6: filters.py:150-507 visualize_filter / visualize_mask and
agent.py:141-204, executed by tests/golden/make_reference_golden.py).  Same OpenCV primitives on the same
inputs -> pixel-exact."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def gold():
  return np.load(os.path.join(HERE, "golden", "reference_golden.npz"))


@pytest.fixture(scope="module")
def filters():
  from exposure_b200 import filters as MF
  from exposure_b200.trainer import default_cfg
  cfg = default_cfg()
  classes = [MF.ExposureFilter, MF.GammaFilter, MF.ImprovedWhiteBalanceFilter, MF.SaturationPlusFilter, MF.ToneFilter,
             MF.ContrastFilter, MF.WNBFilter, MF.ColorFilter, MF.LevelFilter, MF.VignetFilter]
  net = torch.zeros(1, 64, 64, 3)
  return [c(net, cfg) for c in classes]


@pytest.mark.parametrize("fid", range(10))
def test_visualize_filter_and_mask_are_pixel_exact(gold, filters, fid):
  f = filters[fid]
  dbg = {"filter_parameters": np.array(gold["vz%d_param" % fid]), "mask": np.array(gold["vz_mask"])}
  for size in (64, 256):
    canvas = np.full((size, size, 3), 0.5, dtype=np.float32)
    f.visualize_filter(dbg, canvas)
    want = gold["vz%d_canvas%d" % (fid, size)]
    assert canvas.dtype == want.dtype and np.array_equal(canvas, want), (f.get_short_name(), size)
    assert not np.array_equal(want, np.full_like(want, 0.5)) or f.get_short_name() == "V"   # something was drawn
  assert np.array_equal(f.visualize_mask(dbg, (64, 64)), gold["vz%d_maskimg" % fid])


@pytest.mark.parametrize("mode", ["argmax", "sample"])
def test_agent_debugger_is_pixel_exact(gold, filters, mode):
  from exposure_b200.visualize import make_debugger
  p = "vzdbg_%s_" % mode
  host = {"selected_filter_id": int(gold[p + "selected"]), "pdf": np.array(gold[p + "pdf"]),
          "filter_debug_info": [{"filter_parameters": np.array(gold[p + "param%d" % j]), "mask": np.array(gold[p + "mask%d" % j])}
                                for j in range(8)]}
  debugger = make_debugger(filters[:8], 64)
  assert debugger.width == int(gold[p + "width"])
  assert np.array_equal(debugger(host, combined=True), gold[p + "combined"])
  panels = debugger(host, combined=False)
  for got, key in zip(panels, ("panel_pdf", "panel_detail", "panel_mask")):
    assert np.array_equal(got, gold[p + key]), key
