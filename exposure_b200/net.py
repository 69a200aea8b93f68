"""Mirror of net.py's `GAN` driver (net.py:18-877) on top of the explicit train-step schedules:
same entry points (`GAN(cfg, restore)`, `.train()`, `.restore(ckpt)`, `.eval(spec_files, ...)`), same
directory layout (`models/<cfg.name>/model.ckpt-<step>`), same iteration schedule and log line, so
train.py / evaluate.py of the reference run against it with only the import of `net` changed
(INTEGRATION.md).  What it leaves out is outside the hot path: TensorBoard
summaries, the visualisers (net.py:420-670), script back-ups.

cfg is the reference's cfg dict; the keys it must carry beyond trainer.default_cfg():
  name                       model directory under models/
  fake_data_provider / real_data_provider   callables returning objects with get_next_batch(n)
                             -> [n,64,64,3] (the reference's DataProvider interface,
                             data_provider.py; replay.SyntheticProvider when absent)
"""
import os
import statistics
import time

import torch

from . import tf_bundle
from .agent import make_agent_generator
from .critics import make_critic
from .evaluate import evaluate_files
from .replay import ReplayMemory, SyntheticProvider
from .trainer import Trainer, default_cfg


class _DeviceProvider:
  """Adapts a reference-style provider (numpy batches on the host) to device tensors."""

  def __init__(self, inner, device):
    self.inner, self.device = inner, device

  def get_next_batch(self, n):
    out = self.inner.get_next_batch(n)
    img = out[0] if isinstance(out, (tuple, list)) else out          # (images, features) in the reference
    return torch.as_tensor(img, dtype=torch.float32).to(self.device, non_blocking=True)


class GAN:

  def __init__(self, cfg=None, restore=False, device=None, seed=0):
    self.cfg = cfg = cfg or default_cfg()
    assert cfg.get("gan", "w") == "w", "only the WGAN-GP branch is implemented (DESIGN.md section 10)"
    name = cfg.get("name", "b200")
    self.dir = os.path.join("models", name)                        # net.py:27
    os.makedirs(self.dir, exist_ok=True)
    self.device = device or torch.device("cuda", torch.cuda.current_device())
    self.trainer = Trainer(cfg, self.device, seed=seed)
    # cfg.generator / cfg.critic are TF graph builders in the reference (config_example.py:126-127);
    # here they are callables bound to this model's weights with the same signatures
    cfg.generator = make_agent_generator(self.trainer)
    cfg.critic = make_critic(self.trainer)
    cfg.value = make_critic(self.trainer, value=True)
    fake = cfg.get("fake_data_provider")
    real = cfg.get("real_data_provider")
    fake = _DeviceProvider(fake(), self.device) if fake else SyntheticProvider(self.device, "raw", seed + 11)
    real = _DeviceProvider(real(), self.device) if real else SyntheticProvider(self.device, "retouched", seed + 12)
    self.rng = torch.Generator(device=self.device).manual_seed(seed + 13)
    self.memory = ReplayMemory(cfg, fake, real, self.device, seed=seed) if not restore else None   # net.py:43 load=not restore
    if self.memory is not None:
      self.trainer.attach_memory(self.memory, self.rng)
    self.log = []

  # ---- net.py:298-403 ------------------------------------------------------------------------
  def train(self, max_iter_step=None, graphs=True, log_every=10, save_every=500):
    cfg, t = self.cfg, self.trainer
    assert self.memory is not None, "GAN(cfg, restore=True) is for evaluation (no replay memory, net.py:43)"
    steps = cfg.max_iter_step if max_iter_step is None else max_iter_step
    g_pool, v_pool, emd_pool, cgn = [], [], [], 0.0
    start = time.time()
    for it in range(steps + 1):
      it_start = time.time()
      if graphs and it == 1 and hasattr(t, "enable_graphs"):       # iteration 0 runs 100+100 eager steps (net.py:312-322)
        t.enable_graphs(cfg.batch_size)
      out = t.train_iteration(it)
      if it % log_every == 0 or it == steps:                       # the only host synchronisation of the loop
        g_pool.append(float(out["g_loss"])); v_pool.append(float(out["v_loss"]))
        if out["emd"] is not None:
          emd_pool.append(float(out["emd"])); cgn = float(out["critic_gradient_norm"])
        k = cfg.get("median_filter_size", 101)
        g_pool, v_pool, emd_pool = g_pool[-k:], v_pool[-k:], emd_pool[-k:]
        line = "it%6d,%5.0f ms/it, g_loss=%.2f, v_loss=%.2f, EMD=%.3f, cgn=%.2f" % (
            it, 1000 * (time.time() - it_start), statistics.median(g_pool), statistics.median(v_pool),
            statistics.median(emd_pool) if emd_pool else float("nan"), cgn)          # net.py:398-402
        self.log.append(line)
        print(line)
      if (it + 1) % save_every == 0:
        self.save(it + 1)                                          # net.py:383-387
    self.train_seconds = time.time() - start

  def save(self, global_step):
    tf_bundle.save_checkpoint(self.trainer, os.path.join(self.dir, "model.ckpt-%s" % global_step))

  def restore(self, ckpt):                                         # net.py:405-407
    prefix = ckpt if os.path.exists(str(ckpt) + ".index") else os.path.join(self.dir, "model.ckpt-%s" % ckpt)
    tf_bundle.restore_checkpoint(self.trainer, prefix)

  # ---- net.py:711-877 ------------------------------------------------------------------------
  def eval(self, spec_files=None, output_dir="./outputs", step_by_step=False, show_linear=True, show_input=True):
    """Retouch the listed image files (tif: ProPhoto linearisation; png/jpg: sRGB^2.2 / (2 max),
    net.py:726-748) with cfg.test_steps policy steps and write what the reference writes per input:
    `.retouched.png`, `.linear.png`, `.input_tone_mapped.png`, `.intermediateNN.png` (step_by_step),
    `.steps.png` and `_debug.pkl` (exposure_b200/evaluate.py evaluate_files)."""
    return evaluate_files(self.trainer, list(spec_files or []), output_dir=output_dir, generator=self.rng,
                          step_by_step=step_by_step, show_linear=show_linear, show_input=show_input)
