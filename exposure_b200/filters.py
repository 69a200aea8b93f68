"""Drop-in mirror of the reference's `filters.py` Filter registry on the CUDA hot path.

Same class names, constructor `(net, cfg)`, method set and argument meaning as the reference
(filters.py:9-507) so that `cfg.filters = [ExposureFilter, GammaFilter, ...]`
(config_example.py:22-25) works verbatim; tensors are torch CUDA tensors (NHWC float32) and
`process` / `filter_param_regressor` run the kernels of csrc/filters.cu through the C ABI, with
autograd support (ops.FilterProcessFn / FilterRegressFn).  There is no CPU path.

Parameter shapes follow the reference: `[B,1]` (E, G, S+, Ct, BW), `[B,3]` (W),
`[B,1,1,1,8]` (T), `[B,1,1,3,8]` (C)."""
import torch

from . import _cabi
from . import nn_ops as K
from . import ops

__all__ = ["Filter", "ExposureFilter", "GammaFilter", "ImprovedWhiteBalanceFilter", "SaturationPlusFilter",
           "ToneFilter", "ContrastFilter", "WNBFilter", "ColorFilter", "LevelFilter", "VignetFilter", "FILTER_IDS"]


class Filter:
  filter_id = None         # index of the CUDA kernel family (include/exposure_b200.h)

  def __init__(self, net, cfg):
    self.cfg = cfg
    # filters.py:179, 202, 261, 309 read cfg.exposure_range / gamma_range / color_curve_range /
    # tone_curve_range / curve_steps: handed to the library (process-wide, like the reference's one cfg)
    _cabi.set_filter_ranges(cfg)
    self.height, self.width, self.channels = list(map(int, net.shape[1:]))
    self.num_filter_parameters = None
    self.short_name = None
    self.filter_parameters = None
    self.variables = None  # {'fc1/weights','fc1/biases','fc2/weights','fc2/biases'} bound by the owner

  # ---- reference API -------------------------------------------------------------------
  def get_short_name(self):
    assert self.short_name
    return self.short_name

  def get_num_filter_parameters(self):
    assert self.num_filter_parameters
    return self.num_filter_parameters

  def bind_variables(self, variables):
    """The TF reference creates fc1/fc2 under variable_scope('filter_%d') (agent.py:59); here
    the owner of the parameters binds them explicitly."""
    self.variables = variables
    return self

  def extract_parameters(self, features):
    """filters.py:28-44: fc1 (lrelu) -> fc2, split into filter logits and 6 mask logits."""
    assert self.variables is not None, "bind_variables() first"
    v = self.variables
    h = K.fc_fwd(features.contiguous(), v["fc1/weights"], v["fc1/biases"], mode=K.FC_LRELU)
    o = K.fc_fwd(h, v["fc2/weights"], v["fc2/biases"], mode=K.FC_LINEAR)
    n = self.get_num_filter_parameters()
    return o[:, :n], o[:, n:]

  def filter_param_regressor(self, features):
    """features [B, n] raw logits -> regressed parameter in the reference's shape."""
    f = features.reshape(features.shape[0], -1)
    p = ops.FilterRegressFn.apply(f, self.filter_id)[:, :self.get_num_filter_parameters()]
    return self._param_shape(p)

  def _param_shape(self, p):
    return p

  def _flat_params(self, img, param):
    """[Bp, ...] parameter in the reference's shape -> [B, 24]; Bp == 1 broadcasts over the batch like the
    reference's `param[:, None, None, :]` (filters.py:62-99 with specified_parameter), anything else must match."""
    B = img.shape[0]
    flat = param.reshape(param.shape[0], -1)
    if flat.shape[0] != B:
      if flat.shape[0] != 1:
        raise ValueError("parameter batch %d does not match image batch %d" % (flat.shape[0], B))
      flat = flat.expand(B, flat.shape[1])
    if flat.shape[1] > ops.PSTRIDE:
      raise ValueError("at most %d parameters per image, got %d" % (ops.PSTRIDE, flat.shape[1]))
    if flat.shape[1] != ops.PSTRIDE:
      flat = torch.nn.functional.pad(flat, (0, ops.PSTRIDE - flat.shape[1]))
    return flat.contiguous()

  def process(self, img, param):
    """Whole-image filter, no masking (filters.py `process`).  img [B,H,W,3]."""
    return ops.FilterProcessFn.apply(img.contiguous(), self._flat_params(img, param), self.filter_id)

  def debug_info_batched(self):
    return False

  def no_high_res(self):
    return False

  def apply(self, img, img_features=None, specified_parameter=None, high_res=None):
    """filters.py:62-99."""
    assert (img_features is None) ^ (specified_parameter is None)
    if img_features is not None:
      filter_features, mask_parameters = self.extract_parameters(img_features)
      filter_parameters = self.filter_param_regressor(filter_features)
    else:
      assert not self.use_masking()
      filter_parameters = specified_parameter
      mask_parameters = torch.zeros(1, self.get_num_mask_parameters(), device=img.device)
    debug_info = {}
    debug_info["filter_parameters"] = filter_parameters if self.debug_info_batched() else filter_parameters[0]
    self.mask_parameters = mask_parameters
    self.mask = self.get_mask(img, mask_parameters)
    debug_info["mask"] = self.mask[0]
    low_res_output = self._apply_masked(img, filter_parameters, mask_parameters)
    if high_res is not None:
      if self.no_high_res():
        high_res_output = high_res
      else:
        self.high_res_mask = self.get_mask(high_res, mask_parameters)
        high_res_output = self._apply_masked(high_res, filter_parameters, mask_parameters)
    else:
      high_res_output = None
    return low_res_output, high_res_output, debug_info

  def _masks_pixels(self):
    return self.use_masking()

  def _apply_masked(self, img, filter_parameters, mask_parameters):
    """lerp(img, process(img, p), get_mask(img, mask_p)) (filters.py:88).  With masking off the mask
    is ones(1,1,1,1) and lerp(img, processed, 1) == processed for finite img; with masking on the
    mask is evaluated inside the masked step kernel (never materialised)."""
    if not self._masks_pixels():
      return self.process(img, filter_parameters)
    B = img.shape[0]
    flat = self._flat_params(img, filter_parameters)
    ml = mask_parameters.expand(B, mask_parameters.shape[1])
    if ml.shape[1] < ops.MASK_PARAMS:
      ml = torch.nn.functional.pad(ml, (0, ops.MASK_PARAMS - ml.shape[1]))
    return ops.FilterMaskedFn.apply(img.contiguous(), flat, ml.contiguous(), self.filter_id,
                                    float(self.cfg.maximum_sharpness), float(self.cfg.minimum_strength))

  def use_masking(self):
    return self.cfg.masking

  def get_num_mask_parameters(self):
    return 6

  def get_mask(self, img, mask_parameters):
    """filters.py:110-148: mask_parameters are the RAW fc2 outputs [B, 6] (tanh_range inside).
    Returns ones(1,1,1,1) with masking off (every shipped config, config_example.py:36), else the
    [B,H,W,1] mask (debug / visualisation; `apply` evaluates it inside the step kernel instead).
    Not differentiable here -- gradients flow through `apply`."""
    if not self.use_masking():
      return torch.ones(1, 1, 1, 1, device=img.device)
    assert mask_parameters.shape[1] == self.get_num_mask_parameters()
    B = img.shape[0]
    ml = mask_parameters.detach().expand(B, mask_parameters.shape[1])
    if ml.shape[1] < ops.MASK_PARAMS:
      ml = torch.nn.functional.pad(ml, (0, ops.MASK_PARAMS - ml.shape[1]))
    return ops.filter_mask(img.detach().contiguous(), ml.contiguous(), self.filter_id,
                           float(self.cfg.maximum_sharpness), float(self.cfg.minimum_strength), True)

  # ---- visual debugger (host side, cv2; exposure_b200/visualize.py) ------------------------------
  def visualize_filter(self, debug_info, canvas):
    """Draw this filter's parameters onto `canvas` in place (filters.py:150-152 and the per-class
    overrides): debug_info = {'filter_parameters': first image's regressed parameters (numpy), 'mask': ...}."""
    from . import visualize
    visualize.visualize_filter(self.get_short_name(), debug_info, canvas)

  def visualize_mask(self, debug_info, res):
    """filters.py:154-158."""
    from . import visualize
    return visualize.visualize_mask(debug_info, res)

  def draw_high_res_text(self, text, canvas):
    """filters.py:160-168."""
    from . import visualize
    return visualize.draw_high_res_text(text, canvas)


class ExposureFilter(Filter):          # filters.py:170-182
  filter_id = 0

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.short_name = "E"
    self.num_filter_parameters = 1


class GammaFilter(Filter):             # filters.py:194-206
  filter_id = 1

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.short_name = "G"
    self.num_filter_parameters = 1


class ImprovedWhiteBalanceFilter(Filter):   # filters.py:215-238
  filter_id = 2

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.short_name = "W"
    self.channels = 3
    self.num_filter_parameters = self.channels


class SaturationPlusFilter(Filter):    # filters.py:474-498
  filter_id = 3

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.short_name = "S+"
    self.num_filter_parameters = 1


class ToneFilter(Filter):              # filters.py:298-322
  filter_id = 4

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.curve_steps = cfg.curve_steps
    self.short_name = "T"
    self.num_filter_parameters = cfg.curve_steps

  def _param_shape(self, p):
    return p.reshape(-1, 1, self.cfg.curve_steps)[:, None, None, :]


class ContrastFilter(Filter):          # filters.py:404-419
  filter_id = 5

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.short_name = "Ct"
    self.num_filter_parameters = 1


class WNBFilter(Filter):               # filters.py:428-440
  filter_id = 6

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.short_name = "BW"
    self.num_filter_parameters = 1


class ColorFilter(Filter):             # filters.py:247-273
  filter_id = 7

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.curve_steps = cfg.curve_steps
    self.channels = int(net.shape[3])
    self.short_name = "C"
    self.num_filter_parameters = self.channels * cfg.curve_steps

  def _param_shape(self, p):
    return p.reshape(-1, self.channels, self.cfg.curve_steps)[:, None, None, :]


class LevelFilter(Filter):             # filters.py:449-464
  filter_id = 8

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.short_name = "Le"
    self.num_filter_parameters = 2


class VignetFilter(Filter):            # filters.py:341-396
  """process == img * 0 and a mask of its own (5 parameters, quadratic in the grid) that is
  applied whether or not cfg.masking is set -- with masking off the reference multiplies it by 0
  and adds 1, i.e. the output is black (filters.py:390-394)."""
  filter_id = 9

  def __init__(self, net, cfg):
    Filter.__init__(self, net, cfg)
    self.short_name = "V"
    self.num_filter_parameters = 1

  def get_num_mask_parameters(self):
    return 5

  def _masks_pixels(self):
    return True

  def _apply_masked(self, img, filter_parameters, mask_parameters):
    if self.use_masking():
      return Filter._apply_masked(self, img, filter_parameters, mask_parameters)
    return self.process(img, filter_parameters)            # mask*0+1 -> lerp(img, 0, 1)

  def get_mask(self, img, mask_parameters):
    assert mask_parameters.shape[1] == self.get_num_mask_parameters()
    B = img.shape[0]
    ml = torch.nn.functional.pad(mask_parameters.detach().expand(B, 5), (0, 1))
    return ops.filter_mask(img.detach().contiguous(), ml.contiguous(), self.filter_id,
                           float(self.cfg.maximum_sharpness), float(self.cfg.minimum_strength), self.use_masking())


FILTER_IDS = {cls: cls.filter_id for cls in (ExposureFilter, GammaFilter, ImprovedWhiteBalanceFilter,
                                             SaturationPlusFilter, ToneFilter, ContrastFilter, WNBFilter, ColorFilter,
                                             LevelFilter, VignetFilter)}
