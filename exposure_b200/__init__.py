"""exposure_b200 -- B200-native implementation of the Exposure (yuanming-hu/exposure)
data-parallel hot path: the differentiable per-pixel filter stack, the policy/value CNN and
the WGAN-GP critic, behind the reference's Filter-registry / cfg-dict API.

Compute lives in hand-written sm_100a CUDA behind the C ABI of include/exposure_b200.h
(exposure_b200/csrc); this package is the host-side mirror of the reference interface.
There is no CPU fallback."""
__version__ = "0.1.0"
