"""GPU parity of the explicit train-step schedules (exposure_b200/trainer.py) against the
autograd oracle (oracle/train_step.py, fp64 torch CPU; all 8 filters + one-hot select like the
reference).  The CUDA path evaluates only the selected filter and hand-schedules every
backward signal, including the WGAN-GP second-order term as a forward-mode tangent pass.

Tolerance: gradients |err| <= 2e-3 * max|ref| per variable (fp32 kernels vs fp64 oracle through
~12 layers; north_star asks CNN logits <= 1e-3 rel, which is asserted at 1e-4 here)."""
import pytest
import torch

from oracle import filters as OF
from oracle import train_step as OT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trainer(built_lib):
  assert torch.cuda.is_available()
  from exposure_b200.trainer import Trainer
  return Trainer(seed=3)


def _named(trainer, grads=False):
  from exposure_b200.checkpoint import export_named
  d = export_named(trainer, grads=grads)
  return {k: {n: t.double().cpu() for n, t in v.items()} for k, v in d.items()}


def _rel(a, ref):
  return float((a.double().cpu() - ref).abs().max() / (ref.abs().max() + 1e-30))


@pytest.mark.parametrize("masking", [False, True])
def test_generator_step_matches_oracle(trainer, masking):
  """masking=True: cfg.masking (filters.py:62-148) -- the selected filter is blended through its
  spatial mask and the six mask logits of its fc2 head receive gradients."""
  B = 8
  cfg = trainer.cfg
  cfg.masking = masking
  g = torch.Generator().manual_seed(5)
  img = OF.synth_images(B, 64, 64, seed=21, stress=False).double() * 3
  states = torch.zeros(B, 11, dtype=torch.float64)
  states[:, 2] = torch.tensor([0, 1, 2, 3, 4, 4, 6, 7.0])
  states[:, 3:] = (torch.rand(B, 8, generator=g) < 0.3).double()
  noise = torch.rand(B, generator=g, dtype=torch.float64)
  drop_f = (torch.rand(B, 4, 4, 256, generator=g) < 0.5).double() * 2
  drop_s = (torch.rand(B, 4, 4, 256, generator=g) < 0.5).double() * 2
  progress = 0.3
  P = _named(trainer)
  ref = OT.generator_step(P["generator"], P["rl_value"], P["critic"], img, states, noise, drop_f, drop_s, progress, cfg)
  f = lambda t: t.float().cuda().contiguous()
  out = trainer.generator_step(f(img), f(states), f(noise), f(drop_f), f(drop_s), progress, lr_g=1e-5, apply=False)
  assert torch.equal(out["ctx"].ids.cpu(), ref["ids"].to(torch.int32))
  assert _rel(out["fake_output"], ref["fake_output"]) < 1e-5
  assert _rel(out["new_states"], ref["new_states"]) < 1e-6
  assert _rel(out["fake_logit"], ref["fake_logit"]) < 1e-4
  assert _rel(out["old_value"], ref["old_value"]) < 1e-4 and _rel(out["new_value"], ref["new_value"]) < 1e-4
  assert abs(float(out["g_loss"]) - float(ref["g_loss"])) < 1e-4 * (1 + abs(float(ref["g_loss"])))
  assert abs(float(out["v_loss"]) - float(ref["v_loss"])) < 1e-4 * (1 + abs(float(ref["v_loss"])))
  G = _named(trainer, grads=True)
  worst = {}
  for key, refg in (("generator", ref["grads_g"]), ("rl_value", ref["grads_v"])):
    for name, gr in refg.items():
      worst[name] = _rel(G[key][name], gr) if float(gr.abs().max()) > 0 else float(G[key][name].abs().max())
  bad = {k: v for k, v in worst.items() if v > 2e-3}
  cfg.masking = False
  assert not bad, bad
  if masking:    # the mask columns [n, n+6) of the selected filters' fc2 heads are trained
    got = [float(G["generator"]["generator/filter_%d/fc2/weights" % j][:, OF.NUM_PARAMS[j]:].abs().max()) for j in range(8)]
    assert max(got) > 0, got


def test_critic_step_matches_oracle(trainer):
  B = 6
  cfg = trainer.cfg
  g = torch.Generator().manual_seed(8)
  real = (OF.synth_images(B, 64, 64, seed=31, stress=False) * 6).clamp(0, 1.2).double()
  fake = (OF.synth_images(B, 64, 64, seed=32, stress=False) * 3).double()
  alpha = torch.rand(B, generator=g, dtype=torch.float64)
  # scale the critic so that the one-sided penalty is active (||grad|| > 1) for most samples
  with torch.no_grad():
    trainer.cri.p["critic/fully_connected_1/weights"].mul_(40.0)
  P = _named(trainer)
  ref = OT.critic_step(P["critic"], real, fake, alpha, cfg)
  f = lambda t: t.float().cuda().contiguous()
  out = trainer.critic_step(f(real), f(fake), f(alpha), lr_c=1e-5, apply=False)
  assert float(ref["gradient_penalty"]) > 0, "test set-up: penalty inactive"
  assert abs(float(out["emd"]) - float(ref["emd"])) < 1e-4 * (1 + abs(float(ref["emd"])))
  assert abs(float(out["gradient_penalty"]) - float(ref["gradient_penalty"])) < 1e-3 * float(ref["gradient_penalty"])
  assert abs(float(out["critic_gradient_norm"]) - float(ref["critic_gradient_norm"])) < 1e-4 * float(ref["critic_gradient_norm"])
  G = _named(trainer, grads=True)["critic"]
  bad = {n: _rel(G[n], gr) for n, gr in ref["grads_c"].items() if _rel(G[n], gr) > 2e-3}
  assert not bad, bad
  with torch.no_grad():
    trainer.cri.p["critic/fully_connected_1/weights"].mul_(1 / 40.0)


def test_adam_updates_match_oracle(trainer):
  """One applied generator step moves theta_g / theta_v exactly like tf.train.AdamOptimizer."""
  B = 4
  from exposure_b200.trainer import Trainer
  t = Trainer(seed=11)
  g = torch.Generator(device="cuda").manual_seed(1)
  img = torch.rand(B, 64, 64, 3, device="cuda", generator=g) * 0.5
  states = torch.zeros(B, 11, device="cuda")
  noise, df, ds, _ = t.draw(B, generator=g)
  before = t.val.flat.clone()
  out = t.generator_step(img, states, noise, df, ds, 0.0, lr_g=1.5e-5, apply=True)
  gv = t.val.grad.clone()
  p2, _, _ = OT.adam_update(before.double(), gv.double(), torch.zeros_like(gv).double(), torch.zeros_like(gv).double(),
                            10 * 1.5e-5, 1)
  assert _rel(t.val.flat, p2.cpu()) < 1e-6
  assert t.counter_g == 1 and t.counter_v == 1


def test_cuda_graph_replay_matches_eager(built_lib):
  """enable_graphs(): the captured generator / critic steps update the weights exactly like the
  eager launch sequence (same kernels, deterministic reductions), with per-step lr / progress
  fed through device scalars."""
  from exposure_b200.trainer import Trainer
  B = 8
  a, b = Trainer(seed=21), Trainer(seed=21)
  a.cfg.batch_size = b.cfg.batch_size = B
  b.enable_graphs(B)
  assert torch.equal(a.gen.flat, b.gen.flat) and torch.equal(a.cri.flat, b.cri.flat)
  g = torch.Generator(device="cuda").manual_seed(4)
  for it in range(3):
    img = torch.rand(B, 64, 64, 3, device="cuda", generator=g) * 0.6
    states = torch.zeros(B, 11, device="cuda")
    states[:, 2] = it
    noise, df, ds, alpha = a.draw(B, generator=g)
    real = torch.rand(B, 64, 64, 3, device="cuda", generator=g)
    progress, lr_g, lr_c = 0.1 * it, 1e-3 * (it + 1), 2e-3 * (it + 1)
    oa = a.generator_step(img, states, noise, df, ds, progress, lr_g)
    fa = oa["fake_output"].clone()
    ob = b.generator_step(img, states, noise, df, ds, progress, lr_g)
    assert torch.equal(fa, ob["fake_output"])
    ca = a.critic_step(real, fa, alpha, lr_c)
    cb = b.critic_step(real, fa, alpha, lr_c)
    assert torch.allclose(ca["emd"], cb["emd"], rtol=1e-6, atol=1e-7)
  for sa, sb in ((a.gen, b.gen), (a.val, b.val), (a.cri, b.cri)):
    assert torch.allclose(sa.flat, sb.flat, rtol=1e-6, atol=1e-8), float((sa.flat - sb.flat).abs().max())
  assert b.graph_launches["generator"] > 50 and b.graph_launches["critic"] > 30


def test_c_average_moving_average_follows_tf_zero_debias(built_lib):
  """tf.train.ExponentialMovingAverage(decay=0.99, zero_debias=True) of c_average = mean(D(fake) + D(real)) / 2,
  updated once per applied critic step (net.py:119-120, 165-168, 268-269); in the shipped checkpoint
  mul_8/ExponentialMovingAverage/local_step == Variable_2 (critic steps) confirms the cadence."""
  from exposure_b200.trainer import Trainer
  t = Trainer(seed=2)
  g = torch.Generator(device="cuda").manual_seed(4)
  biased, want = 0.0, None
  for k in range(1, 4):
    real = torch.rand(8, 64, 64, 3, device="cuda", generator=g)
    fake = torch.rand(8, 64, 64, 3, device="cuda", generator=g) * 0.3
    alpha = torch.rand(8, device="cuda", generator=g)
    out = t.critic_step(real, fake, alpha, lr_c=1e-5, apply=True)
    c_avg = float(out["c_average"])
    lg = out["logits"]
    assert abs(c_avg - 0.5 * float(lg[:8].mean() + lg[8:16].mean())) < 1e-6
    biased = biased - (biased - c_avg) * (1 - 0.99)
    want = biased / (1 - 0.99 ** k)
    e = t.ema
    assert e["local_step"] == k and abs(e["biased"] - biased) < 1e-5 * (1 + abs(biased)) and abs(e["value"] - want) < 1e-4 * (1 + abs(want))
  before = t.ema
  t.critic_step(real, fake, alpha, lr_c=1e-5, apply=False)        # gradient-only call: opt_c not run, no update
  assert t.ema == before
