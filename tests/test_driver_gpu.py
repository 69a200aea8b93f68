"""GAN driver mirror (exposure_b200/net.py == net.py:18-877): a short training run, checkpoint
round trip through the TF-bundle writer / reader, and the reference-signature callables
cfg.generator / cfg.critic (agent.py:41, critics.py:42)."""
import os

import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]


def test_gan_train_save_restore_eval(built_lib, tmp_path, monkeypatch):
  monkeypatch.chdir(tmp_path)
  from exposure_b200.net import GAN
  from exposure_b200.trainer import default_cfg
  from exposure_b200 import tf_bundle
  cfg = default_cfg()
  cfg.name = "t"
  cfg.batch_size = 16
  cfg.replay_memory_size = 32
  cfg.max_iter_step = 3
  cfg.critic_initialization = 0
  gan = GAN(cfg, seed=3)
  t = gan.trainer
  # keep iteration 0 short: the reference does 100 + 100 warm-up steps there (net.py:312-322);
  # at least test_steps generator steps are needed before the pool holds terminated records
  orig = t.train_iteration
  t.train_iteration = lambda it, **kw: orig(it, giters=14 if it == 0 else 1, citers=2)
  gan.train(graphs=False, log_every=1, save_every=2)
  assert len(gan.log) == 4 and all(np.isfinite(float(l.split("g_loss=")[1].split(",")[0])) for l in gan.log)
  prefix = os.path.join("models", "t", "model.ckpt-2")
  assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
  assert tf_bundle.verify_bundle(prefix) == 240                       # same variable count as the shipped checkpoint
  names = set(tf_bundle.read_index(prefix))
  assert {"generator/Conv/weights", "generator/filter_7/fc2/weights", "critic/fully_connected_1/biases",
          "OptimizeLoss/rl_value/critic/Conv/weights/Adam_1", "OptimizeLoss_1/beta1_power", "Variable_2"} <= names
  gan.save(99)
  snap = [s.flat.clone() for s in (t.gen, t.val, t.cri)] + [s.m.clone() for s in (t.gen, t.val, t.cri)]
  counters = (t.counter_g, t.counter_v, t.counter_c)
  gan2 = GAN(cfg, restore=True, seed=5)
  gan2.restore(99)
  t2 = gan2.trainer
  for a, b in zip(snap, [s.flat for s in (t2.gen, t2.val, t2.cri)] + [s.m for s in (t2.gen, t2.val, t2.cri)]):
    assert torch.equal(a, b)
  assert (t2.counter_g, t2.counter_v, t2.counter_c) == counters
  # reference-signature callables
  B = 4
  img = torch.rand(B, 64, 64, 3, device=t.device) * 0.2
  z = torch.rand(B, cfg.z_dim, device=t.device)
  states = torch.zeros(B, cfg.num_state_dim, device=t.device)
  (net, new_states, surrogate, penalty), info, debugger = cfg.generator(inp=[img, z, states], is_train=0, progress=0.5, cfg=cfg)
  assert net.shape == (B, 64, 64, 3) and new_states.shape == (B, 11) and surrogate.shape == (B, 1) and penalty.shape == (B, 1)
  assert info["selected_filter_ids"].shape == (B,)
  # debug_info of the first image in the reference's structure (agent.py:131-137) + the cv2 debugger (agent.py:141-204)
  assert info["selected_filter_id"] == int(info["selected_filter_ids"][0]) and info["pdf"].shape == (8,)
  assert len(info["filter_debug_info"]) == 8 and info["filter_debug_info"][4]["filter_parameters"].shape == (1, 1, 1, 8)
  assert info["filter_debug_info"][7]["filter_parameters"].shape == (1, 1, 3, 8) and info["filter_debug_info"][0]["mask"].shape == (1, 1, 1)
  sel = info["selected_filter_id"]
  n = [1, 1, 3, 1, 8, 1, 1, 24][sel]
  assert np.allclose(info["filter_debug_info"][sel]["filter_parameters"].reshape(-1),
                     info["selected_filter_parameters"][0, :n].cpu().numpy(), rtol=1e-6, atol=1e-7)
  assert debugger.width == 64
  canvas = debugger(info)
  assert canvas.shape == (64, 64, 3) and canvas.dtype == np.float32 and np.isfinite(canvas).all()
  panels = debugger(info, combined=False)
  assert len(panels) == 3 and all(p.shape == (64, 64, 3) for p in panels)
  logit, _, _ = cfg.critic(images=net, cfg=cfg)
  assert logit.shape == (B, 1) and torch.isfinite(logit).all()
  value, _, _ = cfg.value(images=net, cfg=cfg, states=new_states)
  assert value.shape == (B, 1)
  hi = torch.rand(B, 96, 128, 3, device=t.device) * 0.2
  (net_h, _, hr), _, _ = cfg.generator(inp=[img, z, states], is_train=0, progress=0.5, cfg=cfg, high_res=hi)
  assert hr.shape == hi.shape


def test_gan_eval_writes_the_reference_outputs(built_lib, tmp_path, monkeypatch):
  """GAN.eval (net.py:711-877) per input file: retouched / linear / input_tone_mapped / intermediateNN /
  steps montage / debug pickle; same-resolution files are batched; the step-by-step path ends on the same
  pixels as the single fused full-resolution kernel."""
  import pickle
  import cv2
  monkeypatch.chdir(tmp_path)
  from exposure_b200.net import GAN
  from exposure_b200.trainer import default_cfg
  from exposure_b200.evaluate import evaluate_files
  cfg = default_cfg()
  cfg.name = "e"
  cfg.batch_size = 8
  cfg.replay_memory_size = 16
  gan = GAN(cfg, seed=4)
  rng = np.random.RandomState(0)
  files = []
  for k, (h, w) in enumerate(((90, 120), (90, 120), (70, 70))):
    fn = str(tmp_path / ("in%d.png" % k))
    cv2.imwrite(fn, (rng.rand(h, w, 3) * 255).astype(np.uint8))
    files.append(fn)
  out_dir = str(tmp_path / "outputs")
  gan.rng.manual_seed(11)
  ids = gan.eval(files, output_dir=out_dir, step_by_step=True)
  S = cfg.test_steps
  for fn in files:
    base = os.path.join(out_dir, os.path.basename(fn))
    for suffix in (".retouched.png", ".linear.png", ".input_tone_mapped.png", ".steps.png", "_debug.pkl"):
      assert os.path.exists(base + suffix), suffix
    assert [os.path.exists(base + ".intermediate%02d.png" % s) for s in range(S)] == [True] * (S - 1) + [False]
    steps = cv2.imread(base + ".steps.png")
    assert steps.shape == (68 * 4, 68 * (S + 1), 3)
    infos = pickle.load(open(base + "_debug.pkl", "rb"))
    assert len(infos) == S and [d["selected_filter_id"] for d in infos] == ids[fn]
    assert len(infos[0]["filter_debug_info"]) == 8 and infos[0]["pdf"].shape == (8,)
    ret = cv2.imread(base + ".retouched.png")
    src = cv2.imread(fn)
    assert ret.shape == src.shape
  # the fused single-kernel path (no debug renderings) writes the same retouched pixels for the same draws
  gan.rng.manual_seed(11)
  out2 = str(tmp_path / "outputs2")
  ids2 = evaluate_files(gan.trainer, files, output_dir=out2, generator=gan.rng, show_linear=False, show_input=False, debug=False)
  assert ids2 == ids
  for fn in files:
    a = cv2.imread(os.path.join(out_dir, os.path.basename(fn) + ".retouched.png"))
    b = cv2.imread(os.path.join(out2, os.path.basename(fn) + ".retouched.png"))
    assert np.array_equal(a, b)
    assert not os.path.exists(os.path.join(out2, os.path.basename(fn) + ".steps.png"))
