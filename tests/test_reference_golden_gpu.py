"""GPU parity, direct: the CUDA path (through the C ABI) against vectors produced by executing the
reference's own Python (tests/golden/reference_golden.npz; see tests/golden/make_reference_golden.py
and tests/test_reference_golden.py for what that fixture is and is not).  No oracle in between.

Tolerances (fp32 kernels vs the fixture's fp64 run):
  filter pixels    <= 1e-5 relative (north_star), entries below 1e-2 of the tensor's max measured
                   against that floor; Contrast gets the cosf cancellation allowance of test_filters_gpu.py
  filter gradients <= 1e-4 of max(|ref|, 1e-3 max|ref|)
  CNN logits / values <= 1e-4 (north_star: 1e-3); losses 1e-4; train-step gradients <= 2e-3 of each
                   variable's max (same bar as tests/test_train_step_gpu.py)"""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import seeded_weights  # noqa: E402

pytestmark = pytest.mark.gpu
NUM_PARAMS = [1, 1, 3, 1, 8, 1, 1, 24, 2, 1]
SP, CT, VG = 3, 5, 9


@pytest.fixture(scope="module")
def gold():
  return np.load(os.path.join(HERE, "golden", "reference_golden.npz"))


@pytest.fixture(scope="module")
def ops(built_lib):
  assert torch.cuda.is_available(), "GPU tests need a CUDA device"
  from exposure_b200 import ops as o
  return o


def T(z, k, dtype=torch.float64):
  return torch.from_numpy(np.array(z[k])).to(dtype)


def C(t):
  return t.float().cuda().contiguous()


def err(a, b, floor=1e-3):
  a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
  scale = float(b.abs().max())
  if scale == 0:
    return float(a.abs().max())
  return float(((a - b).abs() / b.abs().clamp_min(floor * scale)).max())


def relmax(a, b):
  a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
  return float((a - b).abs().max() / (b.abs().max() + 1e-30))


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("fid", range(10))
def test_filter_step_matches_reference_code(gold, ops, fid):
  p = "f%d_" % fid
  n = NUM_PARAMS[fid]
  x, lg, gy = T(gold, p + "x"), T(gold, p + "logits"), T(gold, p + "gy")
  params = ops.filter_regress_fwd(C(lg), fid)
  assert err(params[:, :n], T(gold, p + "param").reshape(x.shape[0], -1)) < 2e-6
  y = ops.filter_fwd(C(x), params, fid)
  want = T(gold, p + "y")
  if fid == CT:
    lum = (0.27 * x[..., 0] + 0.67 * x[..., 1] + 0.06 * x[..., 2]).clamp(0, 1)[..., None]
    prm = T(gold, p + "param").reshape(-1, 1, 1, 1).abs()
    tol = 1e-5 * want.abs().clamp_min(1e-4) + prm * x.abs() / (lum + 1e-6) * (2 * 2.0 ** -24) + 1e-5 * 1e-2 * float(want.abs().max())
    assert ((y.cpu().double() - want).abs() <= tol).all()
  elif fid == SP:
    # closed-form ramps vs the reference's fp32 hue round trip: up to 6e-7 * V * p of absolute difference (see
    # tests/test_filters_gpu.py _fwd_tol)
    V = x.clamp(max=1.0).amax(dim=-1, keepdim=True).abs().double()
    prm = T(gold, p + "param").reshape(-1, 1, 1, 1).abs().double()
    tol = 1e-5 * want.abs().clamp_min(1e-2 * float(want.abs().max())) + 1e-6 * V * prm
    assert ((y.cpu().double() - want).abs() <= tol).all()
  else:
    assert err(y, want, floor=1e-2) < 1e-5
  gx, gp = ops.filter_bwd(C(x), C(gy), params, fid)
  want_gx = T(gold, p + "gx")
  if fid == SP:            # away from exact channel ties (see tests/test_reference_golden.py)
    px = x.clamp(max=1.0).reshape(-1, 3)
    keep = ((px[:, 0] != px[:, 1]) & (px[:, 1] != px[:, 2]) & (px[:, 0] != px[:, 2]))
    assert err(gx.cpu().reshape(-1, 3)[keep], want_gx.reshape(-1, 3)[keep]) < 1e-4
  else:
    assert err(gx, want_gx) < 1e-4
  assert err(gp[:, :n], T(gold, p + "gparam").reshape(x.shape[0], -1), floor=1e-2) < 1e-4
  gl = ops.filter_regress_bwd(C(lg), gp, fid)
  assert err(gl[:, :n], T(gold, p + "glogits"), floor=1e-2) < 1e-4


@pytest.mark.parametrize("fid", range(10))
def test_masked_step_matches_reference_code(gold, ops, fid):
  """Filter.apply with cfg.masking: fc2 outputs come from the fixture's own fc weights (fp64 on the host,
  that part is plumbing here), the masked kernels do regressor + process + get_mask + lerp."""
  p = "m%d_" % fid
  n = NUM_PARAMS[fid]
  x, hr, feat, gy = T(gold, p + "x"), T(gold, p + "hr"), T(gold, p + "feat"), T(gold, p + "gy")
  W1, b1, W2, b2 = (T(gold, p + k) for k in ("fc1_weights", "fc1_biases", "fc2_weights", "fc2_biases"))
  h = feat @ W1 + b1
  h = 0.6 * h + 0.4 * h.abs()
  o = h @ W2 + b2
  logits = torch.zeros(x.shape[0], 24, dtype=torch.float64)
  logits[:, :n] = o[:, :n]
  mlog = torch.zeros(x.shape[0], 8, dtype=torch.float64)
  mlog[:, :o.shape[1] - n] = o[:, n:]
  low, mask = ops.filter_masked_fwd(C(x), C(logits), C(mlog), fid, want_mask=True, logits=True)
  assert err(mask, T(gold, p + "mask")) < 1e-5
  assert err(low, T(gold, p + "low"), floor=1e-2) < 2e-5
  high = ops.filter_masked_fwd(C(hr), C(logits), C(mlog), fid, logits=True)
  assert err(high, T(gold, p + "high"), floor=1e-2) < 2e-5
  if fid == SP:
    return
  gx, glog, gmask = ops.filter_masked_bwd(C(x), C(gy), C(logits), C(mlog), fid, logits=True)
  assert err(gx, T(gold, p + "gx")) < 1e-4
  # chain the logit gradients through fc2 on the host and compare with the fixture's fc2 gradients
  go = torch.cat([glog[:, :n].cpu().double(), gmask[:, :o.shape[1] - n].cpu().double()], dim=1)
  assert err(go.sum(dim=0), T(gold, p + "g_fc2_biases"), floor=1e-2) < 2e-4
  assert err(h.t() @ go, T(gold, p + "g_fc2_weights"), floor=1e-2) < 2e-4


# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def trainer(built_lib, gold):
  from exposure_b200.checkpoint import import_named
  from exposure_b200.trainer import Trainer
  t = Trainer(seed=1)
  named = {}
  for which in ("generator", "critic"):
    names = [str(s) for s in gold["seed_%s_varnames" % which]]
    shapes = [tuple(int(v) for v in str(s).split(",")) for s in gold["seed_%s_varshapes" % which]]
    named.update({k: seeded_weights.make(k, s, seed=7) for k, s in zip(names, shapes)})
  import_named(t, named)
  return t


def _named_grads(trainer):
  from exposure_b200.checkpoint import export_named
  d = export_named(trainer, grads=True)
  return {n: g for v in d.values() for n, g in v.items()}


@pytest.mark.parametrize("mode", ["argmax", "sample"])
def test_policy_rollout_matches_reference_code(gold, trainer, mode):
  """5-step agent_generator rollout (agent.py:41-260) with the name-seeded weights; every step starts from
  the fixture's previous output so a decision cannot drift."""
  img = T(gold, "thumbs")
  is_train = 1 if mode == "sample" else 0
  for step in range(5):
    p = "seed_ag_%s_s%d_" % (mode, step)
    c = trainer.generator_forward(C(img), C(T(gold, p + "states")), C(T(gold, p + "z0")), C(T(gold, p + "drop_f") * 2).view(-1, 4, 4, 256),
                                  C(T(gold, p + "drop_s") * 2).view(-1, 4, 4, 256), 0.25, is_train=is_train)
    assert int(c.ids[0]) == int(gold[p + "id0"])
    assert torch.equal(c.new_states.cpu().double(), T(gold, p + "new_states"))
    assert relmax(c.out, T(gold, p + "out")) < 2e-5
    assert err(c.surrogate, T(gold, p + "surrogate"), floor=1e-2) < 1e-4
    assert err(c.penalty, T(gold, p + "penalty"), floor=1e-2) < 1e-4
    img = T(gold, p + "out")


def test_critic_and_value_match_reference_code(gold, trainer):
  img, states = T(gold, "thumbs"), T(gold, "seed_cr_states")
  # untrained (name-seeded) nets give logits that nearly cancel: measured against the batch's largest logit
  assert relmax(trainer.critic.forward(C(img)).logit, T(gold, "seed_cr_logit")) < 1e-4
  assert relmax(trainer.critic.forward(C(img * 2)).logit, T(gold, "seed_cr_logit_x2")) < 1e-4
  assert relmax(trainer.value.forward(C(img), C(states)).logit, T(gold, "seed_cr_value")) < 1e-4


def test_train_steps_match_reference_code(gold, trainer):
  """Generator+value step and critic step (WGAN-GP double backward) vs the losses / gradients the fixture
  built from the reference's generator and critic callables (net.py:92-194 lines restated in the script)."""
  p = "seed_ls_"
  with torch.no_grad():
    trainer.cri.p["critic/fully_connected_1/weights"].mul_(40.0)
  try:
    out = trainer.generator_step(C(T(gold, "thumbs")), C(T(gold, p + "states")), C(T(gold, p + "z0")),
                                 C(T(gold, p + "drop_f") * 2).view(-1, 4, 4, 256), C(T(gold, p + "drop_s") * 2).view(-1, 4, 4, 256),
                                 float(gold[p + "progress"]), lr_g=1e-5, apply=False)
    assert torch.equal(out["new_states"].cpu().double(), T(gold, p + "new_states"))
    assert relmax(out["fake_output"], T(gold, p + "fake_output")) < 2e-5
    for k in ("fake_logit", "old_value", "new_value"):
      assert relmax(out[k], T(gold, p + k)) < 1e-4, k
    for k in ("g_loss", "v_loss"):
      assert abs(float(out[k]) - float(gold[p + k])) < 1e-4 * (1 + abs(float(gold[p + k]))), k
    G = _named_grads(trainer)
    checked = 0
    for k in gold.files:
      if not k.startswith(p + "grad"):
        continue
      kind, name = k[len(p):].split("_", 1)
      name = name.replace(".", "/")
      if kind not in ("grad", "gradsample", "gradnorm") or name.startswith("critic/"):
        continue
      g, want = G[name], T(gold, k)
      if kind == "grad":
        assert relmax(g, want) < 2e-3, name
      elif kind == "gradsample":
        assert float((g.double().cpu().reshape(-1)[::997] - want).abs().max()) < 2e-3 * float(g.abs().max()) + 1e-30, name
      else:
        assert abs(float(g.double().norm()) - float(want)) < 2e-3 * float(want) + 1e-30, name
      checked += 1
    assert checked >= 20
    cr = trainer.critic_step(C(T(gold, p + "real")), C(T(gold, p + "fake_output")), C(T(gold, p + "alpha")), lr_c=1e-5, apply=False)
    assert abs(float(cr["emd"]) - float(gold[p + "emd"])) < 1e-4 * (1 + abs(float(gold[p + "emd"])))
    assert abs(float(cr["gradient_penalty"]) - float(gold[p + "gradient_penalty"])) < 1e-3 * float(gold[p + "gradient_penalty"])
    assert abs(float(cr["critic_gradient_norm"]) - float(gold[p + "critic_gradient_norm"])) < 1e-4 * float(gold[p + "critic_gradient_norm"])
    G = _named_grads(trainer)
    n = 0
    for k in gold.files:
      if not k.startswith(p + "grad"):
        continue
      kind, name = k[len(p):].split("_", 1)
      name = name.replace(".", "/")
      if kind not in ("grad", "gradsample", "gradnorm") or not name.startswith("critic/"):
        continue
      g, want = G[name], T(gold, k)
      if kind == "grad":
        assert relmax(g, want) < 2e-3, name
      elif kind == "gradsample":
        assert float((g.double().cpu().reshape(-1)[::997] - want).abs().max()) < 2e-3 * float(g.abs().max()) + 1e-30, name
      else:
        assert abs(float(g.double().norm()) - float(want)) < 2e-3 * float(want) + 1e-30, name
      n += 1
    assert n >= 12
  finally:
    with torch.no_grad():
      trainer.cri.p["critic/fully_connected_1/weights"].mul_(1 / 40.0)
