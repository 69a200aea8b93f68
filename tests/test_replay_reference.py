"""CPU: exposure_b200.replay.ReplayMemory (pool as device slots + host step/stopped mirrors) draws exactly the
records the reference's ReplayMemory draws (replay_memory.py:64-75, 187-196, 230-273), iteration by iteration,
when both are driven like GAN.train (net.py:325-362) from the same `random` seed.  Expected sequences:
tests/golden/reference_golden.npz `rp_*`, produced by running the reference's replay_memory.py itself
(tests/golden/make_reference_golden.py section 7).  Runs on the CPU device: the selection logic is host code,
the gathers / scatters are plain torch indexing."""
import os

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


class Counting:
  """image k is filled with the value k, like the generating script's stand-in provider"""

  def __init__(self, start, device=torch.device("cpu")):
    self.n = start
    self.device = device

  def get_next_batch(self, bs):
    ids = torch.arange(self.n, self.n + bs, dtype=torch.float32)
    self.n += bs
    return ids[:, None, None, None].expand(bs, 4, 4, 3).contiguous().to(self.device)


@pytest.mark.parametrize("tag", ["a", "b"])
def test_replay_memory_draws_the_reference_sequence(tag):
  _run(tag, torch.device("cpu"))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["a", "b"])
def test_replay_memory_draws_the_reference_sequence_on_the_device(tag):
  """The same fixture with the pool resident in HBM (the product configuration): gathers / scatters on cuda:0."""
  _run(tag, torch.device("cuda", 0))


def _run(tag, device):
  from exposure_b200.replay import ReplayMemory
  from exposure_b200.trainer import default_cfg
  from exposure_b200.util import STATE_REWARD_DIM, STATE_STEP_DIM, STATE_STOPPED_DIM
  gold = np.load(os.path.join(HERE, "golden", "reference_golden.npz"))
  p = "rp_%s_" % tag
  cfg = default_cfg()
  cfg.source_img_size = cfg.real_img_size = 4
  cfg.replay_memory_size, cfg.batch_size = int(gold[p + "pool"]), int(gold[p + "batch"])
  cfg.test_steps = int(gold[p + "test_steps"])
  mem = ReplayMemory(cfg, Counting(0, device), Counting(200000, device), device, seed=int(gold[p + "seed"]))
  B = cfg.batch_size
  kinds, ids, steps = gold[p + "kinds"], gold[p + "ids"], gold[p + "steps"]
  e = 0
  for it in range(40):
    images, states, slots = mem.get_next_fake_batch(B)
    assert kinds[e] == 0
    assert images[:, 0, 0, 0].long().tolist() == ids[e].tolist(), (it, "generator batch")
    assert states[:, STATE_STEP_DIM].int().tolist() == steps[e].tolist()
    e += 1
    new_states = states.clone()                                         # agent.py:208-222
    last = ((states[:, STATE_STEP_DIM] + 1 - cfg.test_steps).abs() < 1e-4).float()
    new_states[:, STATE_REWARD_DIM] = last
    new_states[:, STATE_STOPPED_DIM] = last
    new_states[:, STATE_STEP_DIM] = states[:, STATE_STEP_DIM] + 1
    mem.replace_memory(images, new_states, slots)
    # the host mirrors agree with the device states
    for s in mem.image_pool:
      assert mem.step[s] == int(mem.states[s, STATE_STEP_DIM]) and mem.stopped[s] == bool(mem.states[s, STATE_STOPPED_DIM] > 0)
    if any(mem.stopped[s] for s in mem.image_pool):
      for _ in range(2):
        images, states = mem.replay_fake_batch(B)
        assert kinds[e] == 1
        assert images[:, 0, 0, 0].long().tolist() == ids[e].tolist(), (it, "critic batch")
        assert states[:, STATE_STEP_DIM].int().tolist() == steps[e].tolist()
        assert bool((states[:, STATE_STOPPED_DIM] > 0).all())
        e += 1
  assert e == len(kinds)
