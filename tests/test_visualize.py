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


def test_steps_montage_layout():
  """net.py:843-877: 4 rows x (steps+1) columns of 68-pixel cells; trajectory in row 0, the three debugger
  panels of step i half a cell to the right of image i in rows 1-3 (with the reference's 2 / 4 pixel lifts)."""
  from exposure_b200.evaluate import steps_montage
  S = 3
  traj = [np.full((64, 64, 3), 0.1 * (i + 1), dtype=np.float32) for i in range(S + 1)]
  dec = [np.full((64, 64, 3), 0.5, dtype=np.float32)] * S
  op = [np.full((64, 64, 3), 0.6, dtype=np.float32)] * S
  msk = [np.full((64, 64, 3), 0.7, dtype=np.float32)] * S
  m = steps_montage(traj, dec, op, msk)
  assert m.shape == (272, 272, 3) and m.dtype == np.float32
  for i in range(S + 1):
    assert np.all(m[0:64, 68 * i:68 * i + 64] == np.float32(0.1 * (i + 1)))
  assert np.all(m[0:64, 64:68] == 1.0)                                   # padding stays white
  for i in range(S):
    sx = 68 * i + 34
    assert np.all(m[68:132, sx:sx + 64][:2] == 0.5)                      # rows 68..133 hold the decision, later rows are overdrawn
    assert np.all(m[134:198, sx:sx + 64][:2] == 0.6)
    assert np.all(m[200:264, sx:sx + 64] == 0.7)
  assert np.all(m[:, 68 * S + 34 + 64:][64:] == 1.0)
