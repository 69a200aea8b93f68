"""GPU: the fused multi-step forward (exp_filter_chain_fwd) and the high-resolution eval path."""
import pytest
import torch

from oracle import filters as OF

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(5, 48, 40), (3, 33, 31), (2, 270, 480)])
def test_fused_chain_forward_matches_step_by_step_and_oracle(built_lib, shape):
  from exposure_b200 import ops
  B, H, W = shape
  S = 5
  g = torch.Generator().manual_seed(1)
  x = OF.synth_images(B, H, W, seed=6)
  ids = torch.randint(0, 8, (S, B), generator=g, dtype=torch.int32)
  ids[2, 0] = -1                                     # pdf_sample quirk: black output from that step on
  logits = torch.randn(S, B, 24, generator=g) * 0.7
  y = ops.filter_chain_fwd(x.cuda(), logits.cuda(), ids.cuda(), logits=True).cpu()
  # step by step through the single-step kernels: bit-identical pixels
  cur = x.cuda()
  for s in range(S):
    p = ops.filter_regress_fwd(logits[s].cuda(), ids[s].cuda())
    nxt = ops.filter_fwd(cur, p, ids[s].cuda())
    nxt[ids[s].cuda() < 0] = 0
    cur = nxt
  assert torch.equal(cur.cpu(), y)
  # and the oracle, image by image
  for b in range(B):
    ref = x[b:b + 1]
    for s in range(S):
      f = int(ids[s, b])
      ref = torch.zeros_like(ref) if f < 0 else OF.process(f, ref, OF.regress(f, logits[s, b:b + 1, :OF.NUM_PARAMS[f]]))
    # 5 random filters in sequence amplify the per-step 1e-5 differences (gamma up to x3, contrast's
    # cancellation): bit-equality with the validated single-step kernels is asserted above; against
    # the oracle require 99.9 % of the values within 1e-4 and all within 2e-2
    rel = (y[b] - ref[0]).abs() / ref[0].abs().clamp_min(1e-3)
    assert float((rel <= 1e-4).float().mean()) >= 0.999 and float(rel.max()) <= 2e-2, (b, float(rel.max()))


def test_retouch_high_resolution(built_lib):
  from exposure_b200.evaluate import center_thumbnail, retouch
  from exposure_b200.trainer import Trainer
  t = Trainer(seed=2)
  g = torch.Generator(device="cuda").manual_seed(7)
  hi = torch.exp(torch.randn(3, 333, 500, 3, device="cuda", generator=g) - 3.2).clamp_(0, 4)
  th = center_thumbnail(hi)
  assert th.shape == (3, 64, 64, 3)
  g1 = torch.Generator(device="cuda").manual_seed(9)
  g2 = torch.Generator(device="cuda").manual_seed(9)
  a = retouch(t, hi, generator=g1, fused=True)
  b = retouch(t, hi, generator=g2, fused=False)
  assert a["ids"].shape == (5, 3) and torch.equal(a["ids"], b["ids"])
  assert torch.equal(a["output"], b["output"])
  assert float(a["states"][:, 1].min()) == 1.0       # every trajectory stopped after test_steps
