"""CPU: the C-ABI library builds for sm_100a, loads without a GPU, and exports every symbol
include/exposure_b200.h declares (no compute calls here)."""
import ctypes
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
  names = []
  for fn in sorted(os.listdir(os.path.join(ROOT, "include"))):
    if not fn.endswith(".h"):
      continue
    src = open(os.path.join(ROOT, "include", fn)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names += re.findall(r"\b(exp_[a-z0-9_]+)\s*\(", src)
  return sorted(set(names))


def test_header_declares_something():
  assert len(_declared()) >= 8


def test_library_exports_all_declared_symbols(built_lib):
  lib = ctypes.CDLL(built_lib)
  for name in _declared():
    assert hasattr(lib, name), "missing symbol %s" % name


def test_python_binding_covers_header(built_lib):
  from exposure_b200 import _cabi
  assert sorted(_cabi.SIGNATURES) == _declared()
  l = _cabi.lib()
  assert l.exp_version() == 1
  assert [l.exp_num_filter_params(i) for i in range(10)] == [1, 1, 3, 1, 8, 1, 1, 24, 2, 1]
  assert l.exp_num_filter_params(10) < 0 and b"bad filter id" in l.exp_last_error()
  assert l.exp_filter_bwd_workspace_bytes(0, 1, 1) == 0
  assert l.exp_filter_bwd_workspace_bytes(2, 64, 64) == 65536 * 4 + (2 + 4096) * 32 * 4


def test_argument_validation_needs_no_gpu(built_lib):
  """Bad arguments are rejected before any CUDA call."""
  from exposure_b200 import _cabi
  l = _cabi.lib()
  assert l.exp_filter_fwd(None, None, None, 24, None, 0, 1, 4, 4, 0, None) == -1
  assert l.exp_filter_fwd(16, 16, 16, 24, None, 10, 1, 4, 4, 0, None) == -1         # bad id
  assert l.exp_filter_masked_fwd(16, None, None, 16, 24, 16, 6, None, 0, 1, 4, 4, 1.0, 0.3, 1, 0, None) == -1   # no output
  assert l.exp_filter_masked_fwd(16, 16, None, 16, 24, 16, 5, None, 0, 1, 4, 4, 1.0, 0.3, 1, 0, None) == -1     # mstride < 6
  assert l.exp_filter_masked_bwd(16, 16, 16, 16, 16, 16, 24, 16, 6, None, 0, 1, 4, 4, 1.0, 0.3, 1, 16, 1, 0, None) == -3
  assert l.exp_filter_fwd(16, 16, 16, 2, None, 7, 1, 4, 4, 0, None) == -1           # pstride < 24
  assert l.exp_filter_fwd(16, 16, 16, 24, None, 0, 1, 3, 3, 1, None) == -2          # DIRECT, H*W%4 != 0
  assert l.exp_filter_fwd(20, 16, 16, 24, None, 0, 1, 4, 4, 1, None) == -2          # misaligned
  assert l.exp_filter_bwd(16, 16, 16, 16, 16, 24, None, 0, 1, 4, 4, 16, 1, 0, None) == -3   # workspace


def test_library_is_sm100a_only(built_lib):
  out = subprocess.run(["cuobjdump", "-lelf", built_lib], capture_output=True, text=True).stdout
  archs = set(re.findall(r"sm_\d+a?", out))
  assert archs == {"sm_100a"}, archs
