"""ctypes binding of include/exposure_b200.h.  No torch types cross this boundary.

The product path has NO CPU fallback: if the shared library is missing or a call fails,
an exception is raised (ExposureLibError)."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libexposure_b200.so")

EXP_OK = 0
VARIANT_AUTO, VARIANT_DIRECT, VARIANT_TMA, VARIANT_SCALAR = 0, 1, 2, 3
MAX_FILTER_PARAMS = 24
NUM_FILTERS = 8          # shipped cfg.filters (= number of actions)
NUM_FILTER_KINDS = 10    # + LevelFilter (8), VignetFilter (9)
MASK_PARAMS = 6


class ExposureLibError(RuntimeError):
  pass


_c_void_p, _c_int, _c_size_t, _c_float = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_float

# name -> (restype, argtypes); must list every symbol include/exposure_b200.h declares
SIGNATURES = {
    "exp_version": (_c_int, []),
    "exp_last_error": (ctypes.c_char_p, []),
    "exp_num_filter_params": (_c_int, [_c_int]),
    "exp_set_pdl": (_c_int, [_c_int]),
    "exp_set_filter_ranges": (_c_int, [_c_void_p, _c_int]),
    "exp_get_filter_ranges": (_c_int, [_c_void_p]),
    "exp_filter_regress_fwd": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_void_p]),
    "exp_filter_regress_bwd": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "exp_filter_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_filter_chain_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_filter_chain_fwd_bwd_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "exp_filter_chain_fwd_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int,
                                          _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "exp_filter_chain_fwd_bwd_uniform": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int,
                                                  _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "exp_filter_bwd_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "exp_filter_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int,
                                _c_int, _c_int, _c_int, _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "exp_filter_masked_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_void_p,
                                       _c_int, _c_int, _c_int, _c_int, _c_float, _c_float, _c_int, _c_int, _c_void_p]),
    "exp_filter_masked_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int,
                                       _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, _c_float,
                                       _c_int, _c_void_p, _c_size_t, _c_int, _c_void_p]),
    "exp_set_gemm_backend": (_c_int, [_c_int]),
    "exp_conv_fwd": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, ctypes.c_float, _c_void_p, _c_void_p, _c_void_p,
                              _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_conv_dgrad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int,
                                _c_void_p]),
    "exp_conv_wgrad_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int, _c_int]),
    "exp_conv_wgrad": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, ctypes.c_float, _c_void_p, _c_void_p, _c_int,
                                _c_int, _c_int, _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "exp_fc_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "exp_fc_fwd": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int,
                            _c_int, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "exp_fc_dgrad": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int,
                              _c_int, _c_int, _c_int, _c_void_p]),
    "exp_fc_wgrad": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_conv1_supported": (_c_int, [_c_int, _c_int]),
    "exp_conv1_padded_input_elems": (_c_size_t, [_c_int, _c_int, _c_int]),
    "exp_conv1_pad_input": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, ctypes.c_float, _c_void_p, _c_int, _c_int, _c_int,
                                     _c_void_p]),
    "exp_conv1_pad_weights": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "exp_conv1_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                               _c_int, _c_int, _c_int, _c_void_p]),
    "exp_conv1_wgrad_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int, _c_int]),
    "exp_conv1_wgrad": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _c_void_p,
                                 _c_size_t, _c_void_p]),
    "exp_set_floats": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p]),
    "exp_conv_first_supported": (_c_int, [_c_int, _c_int, _c_int]),
    "exp_conv_first_fwd": (_c_int, [_c_void_p, _c_void_p, _c_int, ctypes.c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                    _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_conv_first_wgrad_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "exp_conv_first_wgrad": (_c_int, [_c_void_p, _c_void_p, _c_int, ctypes.c_float, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                                      _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "exp_conv_first_dgrad": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_conv_enrich32": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_int, ctypes.c_float, _c_void_p, _c_int, _c_int, _c_int,
                                   _c_void_p]),
    "exp_conv_pad_weights32": (_c_int, [_c_void_p, _c_int, _c_int, _c_void_p, _c_void_p]),
    "exp_crc32c": (ctypes.c_uint32, [ctypes.c_uint32, _c_void_p, _c_size_t]),
    "exp_colsum_workspace_bytes": (_c_size_t, [_c_int, _c_int, _c_int]),
    "exp_colsum": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_void_p, _c_size_t, _c_void_p]),
    "exp_stats_fwd": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_stats_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_stats_jvp": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_policy_head_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float,
                                     _c_float, _c_float, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p,
                                     _c_void_p, _c_void_p, _c_void_p]),
    "exp_policy_head_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_float, _c_float,
                                     _c_void_p, _c_void_p, _c_void_p]),
    "exp_overexposure_fwd": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_overexposure_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_rl_losses": (_c_int, [_c_void_p] * 7 + [_c_int, _c_int, _c_float, _c_float, _c_float, _c_float, _c_int, _c_int,
                               _c_void_p, _c_void_p, _c_void_p]),
    "exp_interpolate": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "exp_gp_scale": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_float, _c_int, _c_int, _c_void_p]),
    "exp_critic_inputs": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "exp_critic_scalars": (_c_int, [_c_void_p, _c_void_p, _c_int, _c_float, _c_void_p, _c_float, _c_void_p, _c_void_p]),
    "exp_heads_fc2_fwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int, _c_void_p, _c_int,
                                   _c_void_p, _c_int, _c_int, _c_void_p]),
    "exp_heads_select": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p, _c_int, _c_void_p, _c_int,
                                  _c_void_p]),
    "exp_heads_fc2_bwd": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_int,
                                   _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_void_p, _c_int, _c_void_p]),
    "exp_colsum_multi_workspace_bytes": (_c_size_t, [_c_void_p, _c_void_p, _c_int]),
    "exp_colsum_multi": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_size_t, _c_void_p]),
    "exp_stats_bwd_gin": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_int, _c_int, _c_int, _c_void_p]),
    "exp_replay_ctl_words": (_c_int, []),
    "exp_replay_draw_generator": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, ctypes.c_ulonglong, _c_void_p, _c_void_p, _c_void_p, _c_void_p]),
    "exp_replay_replace": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_float, ctypes.c_ulonglong, _c_void_p, _c_void_p,
                                    _c_void_p, _c_void_p]),
    "exp_replay_draw_critic": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, ctypes.c_ulonglong, _c_void_p, _c_void_p, _c_void_p]),
    "exp_gather_rows": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_int, _c_int, _c_void_p]),
    "exp_train_draws": (_c_int, [ctypes.c_ulonglong, _c_void_p, _c_void_p, _c_int, _c_void_p, _c_size_t, _c_float, _c_void_p]),
    "exp_dp_ipc_handle_bytes": (_c_size_t, []),
    "exp_dp_ipc_export": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
    "exp_dp_ipc_open": (_c_int, [_c_void_p, _c_void_p]),
    "exp_dp_ipc_close": (_c_int, [_c_void_p]),
    "exp_dp_flag_bytes": (_c_size_t, []),
    "exp_dp_max_world": (_c_int, []),
    "exp_dp_allreduce_adam": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_int, _c_int,
                                       _c_void_p, _c_size_t, _c_void_p, _c_size_t, _c_float, _c_float, _c_float, _c_void_p]),
    "exp_adam": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_void_p, _c_float, _c_float, _c_float, _c_float,
                          _c_size_t, _c_void_p]),
}

class FilterRanges(ctypes.Structure):
  """exp_filter_ranges (include/exposure_b200.h)"""
  _fields_ = [("exposure_range", _c_float), ("gamma_range", _c_float), ("tone_lo", _c_float), ("tone_hi", _c_float),
              ("color_lo", _c_float), ("color_hi", _c_float)]


def set_filter_ranges(cfg=None):
  """Hand the regressor ranges of `cfg` (filters.py:179, 202, 261, 309 read cfg.exposure_range, cfg.gamma_range,
  cfg.color_curve_range, cfg.tone_curve_range; cfg.curve_steps) to the library; None restores the defaults.
  Keys a cfg does not carry keep their config_example.py value."""
  if cfg is None:
    check(lib().exp_set_filter_ranges(None, 8), "exp_set_filter_ranges")
    return
  g = lambda k, d: cfg[k] if k in cfg else d
  tone, color = tuple(g("tone_curve_range", (0.5, 2))), tuple(g("color_curve_range", (0.90, 1.10)))
  r = FilterRanges(float(g("exposure_range", 3.5)), float(g("gamma_range", 3)), float(tone[0]), float(tone[1]),
                   float(color[0]), float(color[1]))
  check(lib().exp_set_filter_ranges(ctypes.byref(r), int(g("curve_steps", 8))), "exp_set_filter_ranges")


def get_filter_ranges():
  r = FilterRanges()
  check(lib().exp_get_filter_ranges(ctypes.byref(r)), "exp_get_filter_ranges")
  return {k: getattr(r, k) for k, _ in FilterRanges._fields_}


_lib = None


def lib():
  """The loaded C-ABI library (loads on first use; raises if it was never built)."""
  global _lib
  if _lib is None:
    if not os.path.exists(LIB_PATH):
      raise ExposureLibError(
          "%s not found: run `python -m exposure_b200.build` (there is no CPU fallback)" % LIB_PATH)
    l = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
      fn = getattr(l, name)   # AttributeError if the symbol is missing
      fn.restype = res
      fn.argtypes = args
    _lib = l
    be = os.environ.get("EXPOSURE_GEMM_BACKEND")      # 0 tcgen05 (TMA-fed) where supported, 1 CUDA cores
    if be:
      check(l.exp_set_gemm_backend(int(be)), "exp_set_gemm_backend")
  return _lib


def check(rc, what=""):
  if rc != EXP_OK:
    msg = lib().exp_last_error()
    raise ExposureLibError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))
