"""Committed golden vectors (tests/golden/filters_golden.npz, made by tests/golden/make_golden.py
from the fp64 oracle -- oracle outputs, not reference outputs; those are in test_reference_golden*.py):
  CPU: the oracle still reproduces them (guards the oracle against drift);
  GPU: the CUDA kernels reproduce them through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import filters as F

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filters_golden.npz"))
t64 = lambda k: torch.from_numpy(G[k])


@pytest.mark.parametrize("fid", range(8))
def test_oracle_reproduces_golden(fid):
  x, lg, gy = t64("f%d_x" % fid), t64("f%d_logits" % fid), t64("f%d_gy" % fid)
  p = F.regress(fid, lg)
  assert torch.allclose(p, t64("f%d_params" % fid), rtol=1e-12, atol=1e-14)
  assert torch.allclose(F.process(fid, x, p), t64("f%d_y" % fid), rtol=1e-12, atol=1e-14)
  gx, gp = F.process_bwd_analytic(fid, x, p, gy)
  assert torch.allclose(gx, t64("f%d_gx" % fid), rtol=1e-11, atol=1e-13)
  assert torch.allclose(gp, t64("f%d_gparams" % fid), rtol=1e-11, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("fid", range(8))
def test_cuda_reproduces_golden(built_lib, fid):
  from exposure_b200 import ops
  x, lg, gy = t64("f%d_x" % fid), t64("f%d_logits" % fid), t64("f%d_gy" % fid)
  y_ref, gx_ref, gl_ref = t64("f%d_y" % fid), t64("f%d_gx" % fid), t64("f%d_glogits" % fid)
  xc, lc, gc = x.float().cuda(), lg.float().cuda().contiguous(), gy.float().cuda()
  y = ops.filter_fwd(xc, lc, fid, logits=True).cpu().double()
  tol = 2e-5 * y_ref.abs().clamp_min(1e-4)
  if fid in (F.SP, F.CT):        # the fp32 formula itself is ~1e-4 off fp64 here (cancellation)
    tol = 5e-4 * y_ref.abs().clamp_min(1e-3)
  assert ((y - y_ref).abs() <= tol).all(), float(((y - y_ref).abs() / tol).max())
  gx, gl = ops.filter_bwd(xc, gc, lc, fid, logits=True)
  gx, gl = gx.cpu().double(), gl.cpu().double()
  assert ((gx - gx_ref).abs() <= 2e-4 * gx_ref.abs().clamp_min(1e-3 * float(gx_ref.abs().max()))).all()
  assert torch.allclose(gl, gl_ref, rtol=2e-3, atol=2e-3 * float(gl_ref.abs().max()) + 1e-9)


@pytest.mark.gpu
def test_cuda_chain_reproduces_golden(built_lib):
  from exposure_b200.chain import FilterChain
  ids = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]
  ch = FilterChain(ids)
  lgs = [t64("chain_logits%d" % k).float().cuda().contiguous() for k in range(8)]
  y = ch.forward(t64("chain_x").float().cuda(), lgs).cpu().double()
  gx, gl = ch.backward(t64("chain_gout").float().cuda())
  yr, gr = t64("chain_y"), t64("chain_gx")
  assert ((y - yr).abs() <= 1e-4 * yr.abs().clamp_min(1e-3)).all()
  assert ((gx.cpu().double() - gr).abs() <= 2e-3 * gr.abs().clamp_min(1e-3 * float(gr.abs().max()))).all()
  for k in range(8):
    r = t64("chain_glogits%d" % k)
    assert torch.allclose(gl[k].cpu().double(), r, rtol=5e-3, atol=5e-3 * float(r.abs().max()) + 1e-9), k
