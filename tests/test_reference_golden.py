"""The oracle (oracle/) against vectors produced by executing the REFERENCE'S OWN PYTHON
(tests/golden/reference_golden.npz, made by tests/golden/make_reference_golden.py: the reference's
filters.py / agent.py / critics.py / pdf_sample_layer.py / util.py / net.py (GAN.__init__) run over a TF-1
API stand-in).
This is what pins the oracle: op order, constants, scopes / variable names and formulas are the
reference's; only the TF primitives underneath are restated (see the generating script's header).

Tolerances: the fixture's fp64 run keeps Python-float constants (ln 2, 0.27/0.67/0.06, 0.8, pi, 1/6 ...)
in double where TF -- and the oracle, which mirrors TF -- round them to float32, so fp64-vs-fp64
agreement is bounded by float32 constant rounding (6e-8 relative per constant, more after
cancellation): asserted <= 5e-6 of the tensor's scale, twenty times tighter than anything fp32 can
resolve.  The fixture's fp32 run (the reference's native precision) is compared with the fp32 oracle
at north_star's 1e-5.  S+ image gradients are compared away from exact channel ties only: TF 1.6
registers no RGBToHSV gradient at all, the fixture's come from torch's even-split amax rule, the
oracle / CUDA use first-index arg-max."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import filters as OF
from oracle import nets as ON
from oracle import train_step as OT

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import seeded_weights  # noqa: E402

REF_CKPT = "/root/reference/models/example/pretrained/model.ckpt-20000"


@pytest.fixture(scope="module")
def gold():
  return np.load(os.path.join(HERE, "golden", "reference_golden.npz"))


def T(z, k, dtype=torch.float64):
  return torch.from_numpy(np.array(z[k])).to(dtype)


def err(a, b, floor=1e-3):
  """max |a-b| / max(|b|, floor * max|b|): relative, with tiny entries measured against the tensor's scale."""
  a, b = a.detach().double().reshape(-1), b.detach().double().reshape(-1)
  if b.numel() == 0:
    return 0.0
  scale = float(b.abs().max())
  if scale == 0:
    return float(a.abs().max())
  return float(((a - b).abs() / b.abs().clamp_min(floor * scale)).max())


def test_provenance(gold):
  s = str(gold["provenance"])
  assert "reference Python" in s and "not TensorFlow binaries" in s


# ------------------------------------------------------------------------------------------------
# 1. filter_param_regressor + process + gradients, all ten Filter subclasses
@pytest.mark.parametrize("fid", range(10))
def test_filter_matches_reference_code_fp64(gold, fid):
  p = "f%d_" % fid
  assert str(gold[p + "name"]) == OF.FILTER_NAMES[fid]
  x, lg, gy = T(gold, p + "x"), T(gold, p + "logits"), T(gold, p + "gy")
  param = OF.regress(fid, lg)
  assert err(param, T(gold, p + "param")) < 5e-7
  y = OF.process(fid, x, param)
  assert err(y, T(gold, p + "y")) < 5e-6
  gx, gp = OF.process_bwd_analytic(fid, x, param, gy)
  gl = OF.regress_bwd(fid, lg, gp)
  want_gx = T(gold, p + "gx")
  if fid == OF.SP:
    px = x.clamp(max=1.0).reshape(-1, 3)
    no_tie = ((px[:, 0] != px[:, 1]) & (px[:, 1] != px[:, 2]) & (px[:, 0] != px[:, 2]))
    assert int((~no_tie).sum()) >= 4           # the planted grey / two-channel / black pixels
    gx, want_gx = gx.reshape(-1, 3)[no_tie], want_gx.reshape(-1, 3)[no_tie]
  assert err(gx, want_gx) < 5e-6
  assert err(gp, T(gold, p + "gparam")) < 5e-6
  assert err(gl, T(gold, p + "glogits")) < 5e-6


@pytest.mark.parametrize("fid", range(10))
def test_filter_matches_reference_code_fp32(gold, fid):
  """fp32 oracle vs the reference code run in fp32: north_star's pixel tolerance (1e-5 rel)."""
  p = "f%d_" % fid
  x, lg = T(gold, p + "x", torch.float32), T(gold, p + "logits", torch.float32)
  param = OF.regress(fid, lg)
  assert err(param, T(gold, p + "param32")) < 2e-6
  y = OF.process(fid, x, param)
  tol = 1e-5
  if fid == OF.CT:
    tol = 2e-4      # -cos(pi l)/2 + 1/2 cancels for dark pixels: both fp32 evaluations carry ~1e-7 absolute error there
  assert err(y, T(gold, p + "y32"), floor=1e-2) < tol


# ------------------------------------------------------------------------------------------------
# 2. Filter.apply with masking: extract_parameters (fc1 lrelu, fc2) -> regressor -> get_mask -> lerp
@pytest.mark.parametrize("fid", range(10))
def test_masked_apply_matches_reference_code(gold, fid):
  p = "m%d_" % fid
  names = [str(n) for n in gold[p + "varnames"]]
  assert names == ["filter_%d/fc1/weights" % fid, "filter_%d/fc1/biases" % fid, "filter_%d/fc2/weights" % fid,
                   "filter_%d/fc2/biases" % fid]                                   # scope names == checkpoint names
  n = OF.NUM_PARAMS[fid]
  x, hr, feat, gy = T(gold, p + "x"), T(gold, p + "hr"), T(gold, p + "feat"), T(gold, p + "gy")
  W1, b1, W2, b2 = (T(gold, p + k).requires_grad_(True) for k in ("fc1_weights", "fc1_biases", "fc2_weights", "fc2_biases"))
  for k, t in (("fc1_weights", W1), ("fc2_biases", b2)):                          # the shared initialiser is reproducible
    assert torch.equal(t.detach(), seeded_weights.make("filter_%d/%s" % (fid, k.replace("_", "/")), t.shape, seed=31))
  xr = x.clone().requires_grad_(True)
  o = ON.fc(ON.fc(feat, W1, b1), W2, b2, act=False)
  assert o.shape[1] == n + (5 if fid == OF.VG else 6)
  assert err(o[:, n:], T(gold, p + "mask_parameters")) < 1e-9
  low = OF.apply_masked(fid, xr, o[:, :n], o[:, n:], True)
  assert err(OF.get_mask(fid, x, o[:, n:].detach(), True), T(gold, p + "mask")) < 5e-6
  assert err(low, T(gold, p + "low")) < 5e-6
  high = OF.apply_masked(fid, hr, o[:, :n].detach(), o[:, n:].detach(), True)
  assert err(OF.get_mask(fid, hr, o[:, n:].detach(), True), T(gold, p + "high_mask")) < 5e-6
  assert err(high, T(gold, p + "high")) < 5e-6
  if fid == OF.SP:
    return                        # gradients through RGB<->HSV at planted ties: covered tie-free in section 1
  g = torch.autograd.grad((low * gy).sum(), [xr, W1, b1, W2, b2])
  for got, key in zip(g, ("gx", "g_fc1_weights", "g_fc1_biases", "g_fc2_weights", "g_fc2_biases")):
    assert err(got, T(gold, p + key)) < 2e-5, key


# ------------------------------------------------------------------------------------------------
# weights for sections 3-5
def _seeded(gold, which):
  names = [str(n) for n in gold["seed_%s_varnames" % which]]
  shapes = [tuple(int(v) for v in str(s).split(",")) for s in gold["seed_%s_varshapes" % which]]
  return {n: seeded_weights.make(n, s, seed=7) for n, s in zip(names, shapes)}


def _pretrained():
  if not os.path.exists(REF_CKPT + ".index"):
    pytest.skip("the shipped checkpoint lives in /root/reference (build container only)")
  from exposure_b200 import tf_bundle
  named = tf_bundle.load_bundle(REF_CKPT)
  return {k: torch.from_numpy(np.array(v)).double() for k, v in named.items()
          if np.asarray(v).dtype == np.float32 and np.asarray(v).ndim >= 1}


def _weights(gold, kind):
  if kind == "pre":
    return _pretrained()
  P = _seeded(gold, "generator")
  P.update(_seeded(gold, "critic"))
  return P


def _cfg():
  from exposure_b200.trainer import default_cfg
  return default_cfg()


def _compact(t):
  return t.reshape(t.shape[0], -1)[:, ::37].double()


def _img_view(kind, t):
  return t if kind == "seed" else _compact(t)


def test_seeded_variable_names_are_the_checkpoint_names(gold):
  """The reference code, run over the shim's variable scopes, asks for exactly the variable names / shapes
  of the shipped checkpoint (the generating script asserts the same for the pretrained run)."""
  names = set(str(n) for n in gold["seed_generator_varnames"])
  assert "generator/Conv_3/weights" in names and "generator/action_selection/selector_fc2/biases" in names
  assert "generator/filter_7/fc2/weights" in names and len(names) == 52
  cn = [str(n) for n in gold["seed_critic_varnames"]]
  assert "critic/fully_connected_1/weights" in cn and "rl_value/critic/Conv/weights" in cn and len(cn) == 24
  shapes = dict(zip(cn, [str(s) for s in gold["seed_critic_varshapes"]]))
  assert shapes["critic/Conv/weights"] == "4,4,6,32" and shapes["rl_value/critic/Conv/weights"] == "4,4,17,32"


# 3. agent_generator rollouts
@pytest.mark.parametrize("kind", ["seed", "pre"])
@pytest.mark.parametrize("mode", ["argmax", "sample"])
def test_agent_rollout_matches_reference_code(gold, kind, mode):
  P = _weights(gold, kind)
  cfg = _cfg()
  img = T(gold, "thumbs")
  is_train = 1 if mode == "sample" else 0
  for step in range(5):
    p = "%s_ag_%s_s%d_" % (kind, mode, step)
    states, z0 = T(gold, p + "states"), T(gold, p + "z0")
    drop_f, drop_s = T(gold, p + "drop_f") / 0.5, T(gold, p + "drop_s") / 0.5
    Pg = {k: v.clone().requires_grad_(step == 0) for k, v in P.items() if k.startswith("generator/")}
    out, new_states, surrogate, penalty, ids, pdf = OT.agent_generator(Pg, img, states, z0, drop_f, drop_s, is_train, 0.25, cfg)
    assert int(ids[0]) == int(gold[p + "id0"]), (step, ids)
    assert err(pdf[0], T(gold, p + "pdf0")) < 5e-6
    assert torch.equal(new_states, T(gold, p + "new_states"))
    assert err(surrogate, T(gold, p + "surrogate")) < 5e-6
    assert err(penalty, T(gold, p + "penalty")) < 5e-6
    assert err(_img_view(kind, out), T(gold, p + "out")) < 5e-6
    if step == 0:
      gw = torch.sin(0.37 * torch.arange(out.numel(), dtype=torch.float64)).reshape(out.shape)
      L = (out * gw).sum() + surrogate.sum() * 0.7 + penalty.sum() * 1.3
      keys = [k[len(p) + 5:].replace(".", "/") for k in gold.files if k.startswith(p + "grad_")]
      assert len(keys) == 8
      gs = torch.autograd.grad(L, [Pg[k] for k in keys], allow_unused=True)
      for k, g in zip(keys, gs):
        want = T(gold, p + "grad_" + k.replace("/", "."))
        g = torch.zeros_like(want) if g is None else g
        assert err(g, want) < 2e-5, k
    img = T(gold, p + "out") if kind == "seed" else out.detach()


@pytest.mark.parametrize("kind", ["seed", "pre"])
def test_agent_high_res_branch_matches_reference_code(gold, kind):
  """agent.py:126-129: the selected filter with the low-res parameters applied to the full-res batch."""
  P = _weights(gold, kind)
  cfg = _cfg()
  p = kind + "_ag_hr_"
  img, hr = T(gold, "thumbs"), T(gold, p + "hr")
  states = torch.zeros(img.shape[0], 11, dtype=torch.float64)
  Pg = {k: v for k, v in P.items() if k.startswith("generator/")}
  out, new_states, _, _, ids, _ = OT.agent_generator(Pg, img, states, T(gold, p + "z0"), T(gold, p + "drop_f") / 0.5,
                                                     T(gold, p + "drop_s") / 0.5, 0, 0.0, cfg)
  assert err(_img_view(kind, out), T(gold, p + "out")) < 5e-6
  # oracle restatement of the high-res branch: same per-image logits, selected filter only
  B = img.shape[0]
  w, b = OT._stack(Pg, "generator")
  feat = ON.cnn(ON.enrich(img, states), w, b) * (T(gold, p + "drop_f") / 0.5).reshape(B, -1)
  high = torch.zeros_like(hr)
  for j in range(8):
    h = ON.fc(feat, Pg["generator/filter_%d/fc1/weights" % j], Pg["generator/filter_%d/fc1/biases" % j])
    o = ON.fc(h, Pg["generator/filter_%d/fc2/weights" % j], Pg["generator/filter_%d/fc2/biases" % j], act=False)
    sel = (ids == j).double()[:, None, None, None]
    high = high + sel * OF.apply_filter(j, hr, o[:, :OF.NUM_PARAMS[j]])
  assert err(_img_view(kind, high), T(gold, p + "high")) < 5e-6


# 4. critic / value
@pytest.mark.parametrize("kind", ["seed", "pre"])
def test_critic_and_value_match_reference_code(gold, kind):
  P = _weights(gold, kind)
  p = kind + "_cr_"
  img, states = T(gold, "thumbs"), T(gold, p + "states")
  cp, vp = OT.critic_params(P, "critic"), OT.critic_params(P, "rl_value/critic")
  x = img.clone().requires_grad_(True)
  logit = ON.critic(x, cp)
  assert err(logit, T(gold, p + "logit")) < 5e-6
  assert err(ON.critic(img * 2.0, cp), T(gold, p + "logit_x2")) < 5e-6
  assert err(ON.critic(img, vp, states=states), T(gold, p + "value")) < 5e-6
  (g,) = torch.autograd.grad(logit.sum(), [x])
  assert err(_img_view(kind, g), T(gold, p + "dlogit_dimg"), floor=1e-2) < 1e-4       # stored as float32 for 'seed'
  # the reference code run in its native fp32 vs the fp64 oracle: north_star's CNN tolerance is 1e-3
  assert err(logit, T(gold, p + "logit32")) < 1e-4


# 5. losses and their gradients (generator / value step and critic step with the WGAN-GP double backward)
# src "ls": net.py:92-194 restated line by line in the generating script on top of the reference's callables;
# src "ng": net.py's own GAN.__init__ executed eagerly (placeholders pre-fed), gradients as recorded by its three
# ly.optimize_loss calls.  Same inputs; the two agree to the last bit, and the oracle must match both.
@pytest.mark.parametrize("src", ["ls", "ng"])
@pytest.mark.parametrize("kind", ["seed", "pre"])
def test_train_step_losses_match_reference_code(gold, kind, src):
  P = _weights(gold, kind)
  cfg = _cfg()
  p = kind + "_ls_"                       # inputs and random draws (shared by both sources)
  q = kind + "_" + src + "_"              # expected outputs
  if kind == "seed":
    P["critic/fully_connected_1/weights"] = P["critic/fully_connected_1/weights"] * 40.0
  Pg = {k: v for k, v in P.items() if k.startswith("generator/")}
  Pv = {k: v for k, v in P.items() if k.startswith("rl_value/")}
  Pc = {k: v for k, v in P.items() if k.startswith("critic/")}
  fake_input, real = T(gold, "thumbs"), T(gold, p + "real")
  states = T(gold, p + "states")
  ref = OT.generator_step(Pg, Pv, Pc, fake_input, states, T(gold, p + "z0"), T(gold, p + "drop_f") / 0.5,
                          T(gold, p + "drop_s") / 0.5, float(gold[p + "progress"]), cfg)
  assert torch.equal(ref["new_states"], T(gold, q + "new_states"))
  assert err(_img_view(kind, ref["fake_output"]), T(gold, q + "fake_output")) < 5e-6
  for k in ("fake_logit", "old_value", "new_value", "g_loss", "v_loss"):
    assert err(ref[k], T(gold, q + k)) < 2e-5, k
  checked = 0
  for k in gold.files:
    if not k.startswith(q + "grad"):
      continue
    kind_, name = k[len(q):].split("_", 1)
    name = name.replace(".", "/")
    src = ref["grads_g"] if name.startswith("generator/") else ref["grads_v"] if name.startswith("rl_value/") else None
    if src is None:
      continue
    g = src[name]
    if kind_ == "grad":
      assert err(g, T(gold, k)) < 5e-5, name
    elif kind_ == "gradsample":
      assert err(g.reshape(-1)[::997], T(gold, k)) < 5e-5, name
    else:
      assert abs(float(g.norm()) - float(gold[k])) < 5e-5 * float(gold[k]) + 1e-300, name
    checked += 1
  assert checked >= 8 + 12
  # critic step: fake batch is a fed constant (net.py:362-368)
  fake = T(gold, p + "fake_output") if kind == "seed" else ref["fake_output"]
  cr = OT.critic_step(Pc, real, fake, T(gold, p + "alpha"), cfg)
  assert float(gold[q + "gradient_penalty"]) > 0, "fixture set-up: penalty inactive"
  for k in ("c_loss", "emd", "gradient_penalty", "critic_gradient_norm"):
    assert err(cr[k], T(gold, q + k)) < 2e-5, k
  n = 0
  for name, g in cr["grads_c"].items():
    key = name.replace("/", ".")
    if q + "grad_" + key in gold.files:
      assert err(g, T(gold, q + "grad_" + key)) < 5e-5, name
    else:
      assert err(g.reshape(-1)[::997], T(gold, q + "gradsample_" + key)) < 5e-5, name
      assert abs(float(g.norm()) - float(gold[q + "gradnorm_" + key])) < 5e-5 * float(gold[q + "gradnorm_" + key])
    n += 1
  assert n == 12


def test_net_graph_has_the_checkpoint_variable_counts(gold):
  """GAN.__init__'s tf.get_collection(TRAINABLE_VARIABLES, scope) over the shim's variable store: 52 generator,
  12 value and 12 critic variables (net.py:205-213) -- the shipped checkpoint's non-optimizer tensors."""
  for kind in ("seed", "pre"):
    assert gold[kind + "_ng_n_theta"].tolist() == [52, 12, 12]
    for k in ("g_loss", "v_loss", "c_loss", "emd", "critic_gradient_norm"):
      assert float(gold["%s_ng_%s" % (kind, k)]) == float(gold["%s_ls_%s" % (kind, k)]), k     # restated lines == net.py


def test_pretrained_rollout_behaves_like_the_paper(gold):
  """Layout sanity of the whole stack (HWIO kernels, NHWC flatten, variable scopes, filter order): the shipped
  weights, run through the reference's own agent_generator, retouch the dark RAW thumbnail of A.tif the way
  the paper and SURVEY Appendix B describe -- exposure first (confidently), then saturation / gamma / tone /
  contrast -- and lift its mean from 0.04 to > 0.3.  A transposed kernel or a permuted flatten gives noise."""
  ids = [int(gold["pre_ag_argmax_s%d_id0" % s]) for s in range(5)]
  assert ids[0] == OF.E and float(gold["pre_ag_argmax_s0_pdf0"][OF.E]) > 0.5
  assert len(set(ids)) == 5                                     # filter_usage_penalty: no filter is used twice
  assert set(ids) <= {OF.E, OF.G, OF.SP, OF.T, OF.CT, OF.C, OF.W, OF.BW}
  m0 = float(gold["thumbs"][0].mean())
  m5 = float(gold["pre_ag_argmax_s4_out"][0].mean())
  assert m0 < 0.05 and m5 > 0.3
