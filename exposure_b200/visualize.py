"""Host-side visual debugger (cv2 canvases): what each Filter subclass draws for the step-by-step
renderings (`Filter.visualize_filter / visualize_mask / draw_high_res_text`, filters.py:150-168 and the
per-class overrides at filters.py:184-191, 208-212, 240-244, 275-295, 324-338, 398-401, 421-425,
442-446, 466-471, 500-507) and the three-panel `debugger` closure of agent_generator (agent.py:141-202).

Not on the hot path: pure numpy / OpenCV on the debug_info of ONE image, after the GPU step.  Drawn
pixels are identical to the reference's (same primitives, coordinates, fonts and colours); this module
is table-driven rather than one method per class: a filter is one of four kinds -- a text label, a
colour swatch, a tone curve, a set of colour curves."""
import numpy as np

_FONT_SCALE = 0.3
_BOX = ((8, 40), (56, 52))        # the label / swatch rectangle on a 64x64 canvas
_TEXT_AT = (8, 48)

# short name -> (label format on the 64x64 canvas, long format drawn on larger canvases or None, font scale)
_LABELS = {
    "E": ("EV %+.2f", "Exposure %+.2f", _FONT_SCALE),
    "S+": ("S %+.2f", "Saturation %+.2f", _FONT_SCALE),
    "Ct": ("Ct %+.2f", None, _FONT_SCALE),
    "BW": ("B&W%+.2f", None, _FONT_SCALE),
}


def _cv2():
  import cv2
  return cv2


def _scalar(v):
  return float(np.asarray(v, dtype=np.float64).reshape(-1)[0])


def draw_high_res_text(text, canvas):
  """filters.py:160-168."""
  cv2 = _cv2()
  cv2.putText(canvas, text, (30, 128), cv2.FONT_HERSHEY_SIMPLEX, 0.8, (0, 0, 0), thickness=5)
  return canvas


def _label(canvas, text, scale):
  cv2 = _cv2()
  cv2.rectangle(canvas, _BOX[0], _BOX[1], (1, 1, 1), cv2.FILLED)
  cv2.putText(canvas, text, _TEXT_AT, cv2.FONT_HERSHEY_SIMPLEX, scale, (0, 0, 0))


def _polyline(canvas, weights, color):
  """Normalised cumulative curve through `weights` (one tone / colour curve), segment by segment."""
  cv2 = _cv2()
  height, width = canvas.shape[:2]
  n = len(weights)
  values = np.array([0] + list(weights))
  values /= sum(values) + 1e-30
  for j in range(n):
    values[j + 1] += values[j]
  for j in range(n):
    p1 = tuple(map(int, (width / n * j, height - 1 - values[j] * height)))
    p2 = tuple(map(int, (width / n * (j + 1), height - 1 - values[j + 1] * height)))
    cv2.line(canvas, p1, p2, color, thickness=1)


def visualize_filter(short_name, debug_info, canvas):
  """Draw filter `short_name`'s parameters (debug_info['filter_parameters'], the first image's regressed
  parameters in the reference's shapes) onto `canvas` in place."""
  cv2 = _cv2()
  p = debug_info["filter_parameters"]
  if short_name in _LABELS:
    small, large, scale = _LABELS[short_name]
    if large is not None and canvas.shape[0] != 64:
      draw_high_res_text(large % _scalar(p), canvas)
    else:
      _label(canvas, small % _scalar(p), scale)
  elif short_name == "G":
    _label(canvas, "G 1/%.2f" % (1.0 / _scalar(p)), _FONT_SCALE)
  elif short_name == "Le":
    lo, hi = (float(v) for v in np.asarray(p).reshape(-1)[:2])
    _label(canvas, "%.2f %.2f" % (lo, hi + 1), 0.25)
  elif short_name == "W":
    s = canvas.shape[0]
    cv2.rectangle(canvas, (int(s * 0.2), int(s * 0.4)), (int(s * 0.8), int(s * 0.6)),
                  [float(v) for v in np.asarray(p).reshape(-1)[:3]], cv2.FILLED)
  elif short_name == "V":
    b = _scalar(p)
    cv2.rectangle(canvas, _BOX[0], _BOX[1], (b, b, b), cv2.FILLED)
  elif short_name == "T":
    _polyline(canvas, np.asarray(p).reshape(-1), (0, 0, 0))
  elif short_name == "C":
    curves = np.asarray(p).reshape(3, -1)
    for i in range(3):
      _polyline(canvas, curves[i], tuple(1 if t == i else 0 for t in range(3)))
  else:
    raise ValueError("unknown filter %r" % short_name)


def visualize_mask(debug_info, res):
  """filters.py:154-158: the first image's mask as a 3-channel image of size `res` (nearest neighbour)."""
  cv2 = _cv2()
  mask = np.asarray(debug_info["mask"], dtype=np.float32)
  return cv2.resize(mask * np.ones((1, 1, 3), dtype=np.float32), dsize=res, interpolation=cv2.INTER_NEAREST)


def make_debugger(filters, width):
  """The `debugger(debug_info, combined=True)` closure agent_generator returns (agent.py:141-204).
  filters: the instantiated Filter objects in cfg.filters order.  combined: one 64x64 canvas (mask dimmed
  to 0.8, the selected filter's drawing, the policy's pdf bars); otherwise [pdf panel, detail panel, mask]."""

  def debugger(debug_info, combined=True):
    cv2 = _cv2()
    size, per_col = 8, 4
    sel = int(debug_info["selected_filter_id"])
    pdf = np.asarray(debug_info["pdf"]).reshape(-1)
    shown = [i for i in range(len(filters)) if not pdf[i] < 1e-10]
    assert 0 <= sel < len(filters)
    img = filters[sel].visualize_mask(debug_info["filter_debug_info"][sel], (64, 64)) * 0.8
    panels = [None, None, None]
    if not combined:
      panels[2] = img.copy()
      img = img * 0 + 0.5
    if sel in shown:
      filters[sel].visualize_filter(debug_info["filter_debug_info"][sel], img)
    if not combined:
      panels[1] = img.copy()
      img = img * 0 + 0.5
    for c, i in enumerate(shown):
      x = c // per_col * 30
      y = size * (c % per_col + 1)
      cv2.putText(img, filters[i].get_short_name(), (x + 6, y + 4), cv2.FONT_HERSHEY_SIMPLEX, 0.233, (255, 255, 255))
      bar_w, bar_h = int(pdf[i] * 20), 0.35
      lo = (x + 16, int(y + (1 - bar_h) * size // 2))
      hi = (x + 16 + bar_w, int(y + (1 + bar_h) * size // 2))
      cv2.rectangle(img, (lo[0] - 1, lo[1] - 1), (hi[0] + 1, hi[1] + 1), (1, 1, 1), cv2.FILLED)
      cv2.rectangle(img, lo, hi, (1.0 if i == sel else 0.3, 0.3, 0.3), cv2.FILLED)
    if combined:
      return img
    panels[0] = img.copy()
    return panels

  debugger.width = int(width)
  return debugger
