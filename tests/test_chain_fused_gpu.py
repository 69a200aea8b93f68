"""GPU parity of the whole-chain fused forward+backward (exp_filter_chain_fwd_bwd) against (a) the
per-step kernels it replaces and (b) the CPU oracle.

Tolerances: pixels and image gradients are produced by the same per-pixel device functions as the
per-step kernels -> <= 2e-6 relative to them (separately compiled instantiations may contract FMAs
differently); parameter gradients differ by reduction order only -> <= 1e-5 of max(|ref|, 1e-3 scale).
Against the fp64 oracle: the per-step tests' bars (pixels 1e-5 through 8 steps -> 2e-4, gradients 2e-3)."""
import pytest
import torch

from oracle import filters as F

pytestmark = pytest.mark.gpu
CHAIN = [F.E, F.G, F.W, F.SP, F.T, F.CT, F.BW, F.C]


@pytest.fixture(scope="module")
def ops(built_lib):
  assert torch.cuda.is_available()
  from exposure_b200 import ops as o
  return o


def _rel(a, b, floor=1e-3):
  a, b = a.detach().double().cpu(), b.detach().double().cpu()
  return float(((a - b).abs() / b.abs().clamp_min(floor * float(b.abs().max()) + 1e-30)).max())


def _run_both(ids, B, H, W, seed, scale=0.5):
  from exposure_b200.chain import FilterChain, FusedFilterChain
  x = F.synth_images(B, H, W, seed=seed)
  lgs = [F.synth_logits(f, B, seed=seed + 1 + k) * scale for k, f in enumerate(ids)]
  gout = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed + 50))
  ref = FilterChain(ids)
  y0 = ref.forward(x.cuda(), [l.cuda() for l in lgs]).clone()
  gx0, gl0 = ref.backward(gout.cuda())
  gx0, gl0 = gx0.clone(), [g.clone() for g in gl0]
  fused = FusedFilterChain(ids, B, torch.device("cuda"))
  fused.set_logits([l.cuda() for l in lgs])
  y1, gx1, _ = fused.forward_backward(x.cuda(), gout.cuda())
  torch.cuda.synchronize()
  return x, lgs, gout, (y0, gx0, gl0), (y1, gx1, fused.glogits_list())


@pytest.mark.parametrize("shape", [(4, 64, 64), (2, 33, 31), (3, 1, 1), (1, 7, 5), (2, 128, 96), (1, 512, 640)])
def test_fused_chain_matches_per_step_kernels(ops, shape):
  B, H, W = shape
  _, _, _, (y0, gx0, gl0), (y1, gx1, gl1) = _run_both(CHAIN, B, H, W, seed=7)
  assert _rel(y1, y0) < 2e-6
  assert _rel(gx1, gx0) < 2e-6
  for k, (a, b) in enumerate(zip(gl1, gl0)):
    assert _rel(a, b[:, :a.shape[1]], floor=1e-2) < 1e-5, k


def test_fused_chain_matches_oracle(ops):
  B, H, W = 3, 48, 40
  x, lgs, gout, _, (y1, gx1, gl1) = _run_both(CHAIN, B, H, W, seed=21)
  y64, gx64, glg64 = F.chain_fwd_bwd(CHAIN, x.double(), [l.double() for l in lgs], gout.double())
  assert _rel(y1, y64) < 2e-4
  assert _rel(gx1, gx64) < 2e-3
  for k, (a, b) in enumerate(zip(gl1, glg64)):
    assert _rel(a, b, floor=1e-2) < 2e-3, k


def test_fused_chain_per_image_ids_short_chains_and_black_step(ops):
  """Per-image filter ids (each image its own episode), S < 8, repeated filters, and id -1 (the pdf_sample
  u == 0 quirk: black output, zero gradients)."""
  from exposure_b200.chain import FusedFilterChain
  B, H, W = 5, 32, 24
  g = torch.Generator().manual_seed(3)
  for S in (1, 3, 5):
    ids = torch.randint(0, 10, (S, B), generator=g).to(torch.int32)
    x = F.synth_images(B, H, W, seed=40 + S)
    gout = torch.randn(B, H, W, 3, generator=g)
    logits = torch.randn(S, B, 24, generator=g) * 0.5
    fused = FusedFilterChain([ids[s].cuda() for s in range(S)], B, torch.device("cuda"))
    fused.logits.copy_(logits.cuda())
    y, gx, gl = fused.forward_backward(x.cuda(), gout.cuda())
    # reference: per-step kernels with per-image ids
    acts, cur = [x.cuda()], x.cuda()
    for s in range(S):
      cur = ops.filter_fwd(cur, fused.logits[s].contiguous(), ids[s].cuda(), logits=True)
      acts.append(cur)
    assert _rel(y, acts[-1]) < 2e-6
    gcur = gout.cuda()
    for s in reversed(range(S)):
      gcur, gp = ops.filter_bwd(acts[s], gcur, fused.logits[s].contiguous(), ids[s].cuda(), logits=True)
      assert _rel(gl[s], gp, floor=1e-2) < 1e-5, (S, s)
    assert _rel(gx, gcur) < 2e-6
  # a black step in the middle
  ids = torch.tensor([[0, 1, 2, 3, 4], [-1, 5, -1, 6, 7], [4, 4, 4, 4, 4]], dtype=torch.int32)
  fused = FusedFilterChain([ids[s].cuda() for s in range(3)], B, torch.device("cuda"))
  fused.logits.copy_((torch.randn(3, B, 24, generator=g) * 0.5).cuda())
  x = F.synth_images(B, H, W, seed=77)
  gout = torch.randn(B, H, W, 3, generator=g)
  y, gx, gl = fused.forward_backward(x.cuda(), gout.cuda())
  assert float(gx[0].abs().max()) == 0 and float(gx[2].abs().max()) == 0 and float(gx[1].abs().max()) > 0
  assert float(gl[0][0].abs().max()) == 0 and float(gl[1][0].abs().max()) == 0       # nothing flows through a black step
  zero = torch.zeros(1, H, W, 3, device="cuda")
  want0 = ops.filter_fwd(zero, fused.logits[2][0:1].contiguous(), 4, logits=True)
  assert _rel(y[0:1], want0) < 2e-6 or float((y[0:1] - want0).abs().max()) == 0


def test_fused_chain_outputs_optional_and_graph(ops):
  from exposure_b200.chain import FusedFilterChain
  B, H, W = 4, 64, 64
  x = F.synth_images(B, H, W, seed=5).cuda()
  gout = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(6)).cuda()
  fused = FusedFilterChain(CHAIN, B, torch.device("cuda"))
  fused.set_logits([(F.synth_logits(f, B, seed=9) * 0.5).cuda() for f in CHAIN])
  y, gx, gl = fused.forward_backward(x, gout)
  y, gx, gl = y.clone(), gx.clone(), gl.clone()
  y2, gx2, gl2 = fused.forward_backward(x, gout, need_output=False, need_input_grad=False)
  assert y2 is None and gx2 is None and torch.equal(gl2, gl)                          # deterministic reduction
  yg, gxg, glg = fused.capture(x, gout)
  fused.replay()
  torch.cuda.synchronize()
  assert torch.equal(yg, y) and torch.equal(gxg, gx) and torch.equal(glg, gl)


def test_fused_chain_rejects_bad_arguments(ops):
  x = torch.zeros(2, 8, 8, 3, device="cuda")
  ids = torch.zeros(9, 2, dtype=torch.int32, device="cuda")
  with pytest.raises(RuntimeError):
    ops.filter_chain_fwd_bwd(x, x, torch.zeros(9, 2, 24, device="cuda"), ids)          # S > 8
  with pytest.raises(ValueError):
    ops.filter_chain_fwd_bwd(x, x, torch.zeros(2, 2, 8, device="cuda"), ids[:2])       # wrong params shape


def test_host_pipelined_fused_chain_matches_resident(ops):
  """HostPipelinedChain(fused=True): host buffers in, host buffers out, one kernel per sub-batch."""
  from exposure_b200.chain import FilterChain, HostPipelinedChain
  B, H, W = 8, 64, 64
  x = F.synth_images(B, H, W, seed=31)
  lgs = [(F.synth_logits(f, B, seed=7) * 0.5).cuda() for f in CHAIN]
  gout = torch.randn(B, H, W, 3, device="cuda")
  ref = FilterChain(CHAIN)
  y = ref.forward(x.cuda(), lgs).clone()
  _, gl = ref.backward(gout)
  hx = x.pin_memory()
  hy = torch.empty(B, H, W, 3).pin_memory()
  hg = [torch.empty(B, F.NUM_PARAMS[f]).pin_memory() for f in CHAIN]
  pipe = HostPipelinedChain(CHAIN, B, H, W, torch.device("cuda"), chunks=4, fused=True)
  for wait in (True, True, False, False):
    hy.zero_()
    pipe.step(hx, lgs, gout, hy, hg, wait=wait)
    pipe.wait()
    assert _rel(hy, y) < 2e-6
    for a, b in zip(hg, gl):
      assert _rel(a, b, floor=1e-2) < 1e-5


def test_fused_chain_full_size_properties(ops):
  """BASELINE configs[1] full size (64x512x512x3), size-independent properties: pixels / image gradients
  bit-identical to the 16 per-step kernels; the backward is exactly linear in dL/dy under power-of-two
  scaling; parameter gradients of an image equal the sum over its two halves; launches are deterministic."""
  from exposure_b200.chain import FilterChain, FusedFilterChain
  B, H, W = 64, 512, 512
  g = torch.Generator(device="cuda").manual_seed(2)
  x = torch.exp(torch.randn(B, H, W, 3, device="cuda", generator=g) - 3.2).clamp_(0, 4)
  gout = torch.randn(B, H, W, 3, device="cuda", generator=g)
  lgs = [torch.randn(B, F.NUM_PARAMS[f], device="cuda", generator=g) * 0.5 for f in CHAIN]
  ref = FilterChain(CHAIN)
  y0 = ref.forward(x, lgs)
  gx0, gl0 = ref.backward(gout)
  fused = FusedFilterChain(CHAIN, B, torch.device("cuda"))
  fused.set_logits(lgs)
  y1, gx1, gl1 = fused.forward_backward(x, gout)
  assert torch.equal(y1, y0) and torch.equal(gx1, gx0)
  for k, (a, b) in enumerate(zip(fused.glogits_list(), gl0)):
    assert _rel(a, b, floor=1e-2) < 5e-5, k     # reduction order only (per-thread sums over the tile loop)
  gl1 = gl1.clone()
  del ref, y0, gx0
  _, gx2, gl2 = fused.forward_backward(x, gout * 2.0, need_output=False)
  assert torch.equal(gx2, gx1 * 2.0)
  assert _rel(gl2, gl1 * 2.0, floor=1e-3) < 1e-5
  _, _, gl3 = fused.forward_backward(x, gout, need_output=False, need_input_grad=False)
  assert torch.equal(gl3, gl1)
  # halves: the regressor chain rule is linear in the parameter gradient, so dL/dlogits adds up as well
  top = FusedFilterChain(CHAIN, B, torch.device("cuda"))
  top.logits.copy_(fused.logits)
  _, _, ga = top.forward_backward(x[:, :256].contiguous(), gout[:, :256].contiguous(), need_output=False, need_input_grad=False)
  ga = ga.clone()
  _, _, gb = top.forward_backward(x[:, 256:].contiguous(), gout[:, 256:].contiguous(), need_output=False, need_input_grad=False)
  assert _rel(ga + gb, gl1, floor=1e-2) < 1e-4


def test_filter_chain_autograd_node(ops):
  """ops.filter_chain under torch autograd == composing the per-step autograd nodes."""
  B, H, W, S = 3, 40, 36, 4
  g = torch.Generator().manual_seed(12)
  ids = torch.tensor([[0, 4, 7], [1, 5, 2], [3, 6, 4], [7, 0, 1]], dtype=torch.int32).cuda()
  x = F.synth_images(B, H, W, seed=13).cuda().requires_grad_(True)
  logits = (torch.randn(S, B, 24, generator=g) * 0.5).cuda().requires_grad_(True)
  w = torch.randn(B, H, W, 3, generator=g).cuda()
  y = ops.filter_chain(x, logits, ids)
  (y * w).sum().backward()
  gx, gl = x.grad.clone(), logits.grad.clone()
  x2 = x.detach().clone().requires_grad_(True)
  l2 = logits.detach().clone().requires_grad_(True)
  cur = x2
  for s in range(S):
    p = ops.FilterRegressFn.apply(l2[s], ids[s])
    cur = ops.FilterProcessFn.apply(cur, p, ids[s])
  (cur * w).sum().backward()
  assert _rel(y, cur) < 2e-6 and _rel(gx, x2.grad) < 2e-6
  for s in range(S):
    for b in range(B):
      n = F.NUM_PARAMS[int(ids[s, b])]
      assert _rel(gl[s, b, :n], l2.grad[s, b, :n], floor=1e-2) < 2e-5, (s, b)


@pytest.mark.parametrize("shape", [(4, 64, 64), (2, 128, 96), (3, 512, 512), (1, 360, 644)])
def test_compile_time_chain_equals_run_time_chain(ops, shape):
  """exp_filter_chain_fwd_bwd_uniform: the kernel specialised for the cfg.filters order E,G,W,S+,T,Ct,BW,C
  (template pack, accumulators in registers / per-thread shared memory, scale steps recomputed, Gamma's backward
  fed with the forward output) against the run-time kernel on the same chain: y and dL/dx BIT-IDENTICAL, parameter
  gradients equal to reduction order; and against the per-image-ids entry point."""
  from exposure_b200.chain import FusedFilterChain
  B, H, W = shape
  x = F.synth_images(B, H, W, seed=91).cuda()
  lgs = [(F.synth_logits(f, B, seed=92 + k) * 0.7).cuda() for k, f in enumerate(CHAIN)]
  gout = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(93)).cuda()
  outs = []
  for ids, static in ((CHAIN, True), (CHAIN, False), ([torch.full((B,), f, dtype=torch.int32, device="cuda") for f in CHAIN], True)):
    fz = FusedFilterChain(ids, B, torch.device("cuda"), static_chain=static)
    fz.set_logits(lgs)
    y, gx, _ = fz.forward_backward(x, gout)
    torch.cuda.synchronize()
    outs.append((y.clone(), gx.clone(), [g.clone() for g in fz.glogits_list()]))
  (ys, gxs, gls), (yr, gxr, glr), (yt, gxt, glt) = outs
  assert torch.equal(ys, yr) and torch.equal(gxs, gxr), "compile-time chain differs from the run-time kernel"
  assert torch.equal(yr, yt) and torch.equal(gxr, gxt)
  for k, (a, b) in enumerate(zip(gls, glr)):
    assert _rel(a, b, floor=1e-2) < 1e-5, k
  # optional outputs
  fz = FusedFilterChain(CHAIN, B, torch.device("cuda"))
  fz.set_logits(lgs)
  y, gx, _ = fz.forward_backward(x, gout, need_output=False)
  assert y is None and torch.equal(gx, gxs)
  y, gx, _ = fz.forward_backward(x, gout, need_input_grad=False)
  assert gx is None and torch.equal(y, ys)


def test_compile_time_chain_repeated_launches_are_reproducible(ops):
  """Deterministic reduction: two launches give bitwise equal parameter gradients; the workspace is left clean."""
  from exposure_b200.chain import FusedFilterChain
  B, H, W = 5, 256, 256
  x = F.synth_images(B, H, W, seed=95).cuda()
  lgs = [(F.synth_logits(f, B, seed=96 + k) * 0.7).cuda() for k, f in enumerate(CHAIN)]
  gout = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(97)).cuda()
  fz = FusedFilterChain(CHAIN, B, torch.device("cuda"))
  fz.set_logits(lgs)
  fz.forward_backward(x, gout)
  g1 = fz.glogits.clone()
  fz.forward_backward(x, gout)
  torch.cuda.synchronize()
  assert torch.equal(g1, fz.glogits)
