"""GPU: the device replay memory (csrc/replay.cu) -- kernels vs the host build of the same logic, pool invariants over
many iterations, and the whole train iteration replayed from ONE CUDA graph."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
S = 11


@pytest.fixture(scope="module")
def rl(tmp_path_factory):
  out = str(tmp_path_factory.mktemp("rl") / "libreplay_harness.so")
  subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", out, os.path.join(HERE, "host_math", "replay_harness.cpp")], check=True)
  l = ctypes.CDLL(out)
  l.rl_draw_generator.restype = ctypes.c_int
  l.rl_draw_critic.restype = ctypes.c_int
  return l


def _p(a):
  return a.ctypes.data_as(ctypes.c_void_p)


def test_kernels_equal_the_host_build_of_the_logic(built_lib, rl):
  from exposure_b200 import _cabi
  from exposure_b200 import nn_ops as K
  P, B, seed = 128, 64, 987654321
  g = torch.Generator().manual_seed(1)
  states = torch.zeros(P, S)
  states[:, 2] = torch.randint(0, 8, (P,), generator=g).float()
  states[:, 1] = (torch.rand(P, generator=g) < 0.25).float()
  ctl = torch.zeros(_cabi.lib().exp_replay_ctl_words(), dtype=torch.int32, device="cuda")
  batch = torch.zeros(B, dtype=torch.int64, device="cuda")
  rest = torch.zeros(P, dtype=torch.int32, device="cuda")
  ds = states.cuda()
  K.replay_draw_generator(ds, B, seed, ctl, batch, rest)
  hb, hr, fu = np.zeros(B, np.int64), np.zeros(P, np.int32), ctypes.c_int()
  sn = states.numpy()
  n_rest = rl.rl_draw_generator(_p(sn), S, P, B, ctypes.c_ulonglong(seed), ctypes.c_ulonglong(0), _p(hb), _p(hr), ctypes.byref(fu))
  c = ctl.cpu().tolist()
  assert batch.cpu().tolist() == hb.tolist() and c[2] == n_rest and c[3] == fu.value and c[0] == 1
  assert rest.cpu()[:n_rest].tolist() == hr[:n_rest].tolist()
  new_states = torch.zeros(B, S)
  new_states[:, 2] = torch.randint(1, 9, (B,), generator=g).float()
  src = torch.zeros(P, dtype=torch.int64, device="cuda")
  K.replay_replace(new_states.cuda(), P, 7, 0.5, seed, ctl, rest, src)
  hs = np.zeros(P, np.int64)
  rl.rl_replace(_p(new_states.numpy()), S, P, B, 7, ctypes.c_float(0.5), ctypes.c_ulonglong(seed), ctypes.c_ulonglong(1), _p(hr), n_rest,
                fu.value, _p(hs))
  assert src.cpu().tolist() == hs.tolist() and int(ctl[0]) == 2
  cb = torch.zeros(B, dtype=torch.int64, device="cuda")
  K.replay_draw_critic(ds, B, seed, ctl, cb)
  hc = np.zeros(B, np.int64)
  nt = rl.rl_draw_critic(_p(sn), S, P, B, ctypes.c_ulonglong(seed), ctypes.c_ulonglong(2), _p(hc))
  assert cb.cpu().tolist() == hc.tolist() and int(ctl[5]) == nt and int(ctl[4]) == 0
  K.replay_draw_critic(torch.zeros(P, S, device="cuda"), B, seed, ctl, cb)        # no terminated record: error flag
  assert int(ctl[4]) == 1
  # gather + draws
  src_rows = torch.randn(300, 64, 64, 3, device="cuda")
  idx = torch.randint(0, 300, (B,), device="cuda")
  out = torch.empty(B, 64, 64, 3, device="cuda")
  assert torch.equal(K.gather_rows(src_rows, idx, out), src_rows[idx])
  st = torch.randn(300, S, device="cuda")
  o2 = torch.empty(B, S, device="cuda")
  assert torch.equal(K.gather_rows(st, idx, o2), st[idx])
  uni, mask = torch.zeros(5000, device="cuda"), torch.zeros(2, 64, 4096, device="cuda")
  before = int(ctl[0])
  K.train_draws(seed, ctl, uniform=uni, mask=mask, keep=0.5)
  assert int(ctl[0]) == before + 1
  assert 0 <= float(uni.min()) and float(uni.max()) < 1 and abs(float(uni.mean()) - 0.5) < 0.02
  assert set(mask.unique().tolist()) == {0.0, 2.0} and abs(float(mask.mean()) - 1.0) < 0.01      # floor(0.5 + U) / 0.5
  u2 = torch.zeros_like(uni)
  K.train_draws(seed, ctl, uniform=u2)
  assert float((u2 != uni).float().mean()) > 0.99                                              # next call, new numbers


def test_device_memory_invariants_over_iterations(built_lib):
  from exposure_b200.replay import DeviceReplayMemory, SyntheticProvider
  from exposure_b200.trainer import default_cfg
  cfg = default_cfg()
  dev = torch.device("cuda", 0)
  mem = DeviceReplayMemory(cfg, SyntheticProvider(dev, "raw", 1), SyntheticProvider(dev, "real", 2), dev, seed=4)
  P, B = mem.P, mem.B
  steps_seen = []
  for it in range(60):
    img, st, _ = mem.get_next_fake_batch(B)
    assert not bool((st[:, 1] > 0).any()), "a terminated record entered a generator batch"
    new = st.clone()                                                      # agent.py:208-222
    last = ((st[:, 2] + 1 - cfg.test_steps).abs() < 1e-4).float()
    new[:, 0], new[:, 1], new[:, 2] = last, last, st[:, 2] + 1
    mem.replace_memory(img * 0.5, new)
    pool_states = mem.states[:P]
    steps_seen.append(float(pool_states[:, 2].mean()))
    if bool((pool_states[:, 1] > 0).any()):
      fake, _ = mem.replay_fake_batch(B)
      src = mem.crit_src.cpu()
      assert bool((mem.states[src, 1] > 0).all())
      assert torch.equal(fake, mem.images[src.cuda()])
  mem.check()
  assert 0.5 < steps_seen[-1] < 5.0 and float(mem.states[:P, 2].max()) <= cfg.maximum_trajectory_length + 1
  assert float(mem.states[P + B:].abs().max()) == 0.0                     # fresh records start from the initial state


def test_whole_iteration_as_one_graph(built_lib):
  """Trainer.enable_iteration_graph: one replay = draw + generator step + re-insert + 5 x (draw + critic step); the
  parameters move, stay finite, the replay-memory counters advance, and the run is reproducible for a seed."""
  from exposure_b200.replay import DeviceReplayMemory, SyntheticProvider
  from exposure_b200.trainer import Trainer, default_cfg

  def run():
    dev = torch.device("cuda", 0)
    cfg = default_cfg()
    cfg.batch_size, cfg.replay_memory_size = 16, 32
    t = Trainer(cfg, dev, seed=0)
    mem = DeviceReplayMemory(cfg, SyntheticProvider(dev, "raw", 10), SyntheticProvider(dev, "real", 20), dev, seed=3)
    t.attach_memory(mem, torch.Generator(device=dev).manual_seed(30))
    t.train_iteration(0, giters=12, citers=1)
    t.enable_iteration_graph()
    p0 = t.gv.flat.clone()
    calls0 = int(mem.ctl[0])
    outs = [t.train_iteration(it, giters=1, citers=5) for it in range(1, 5)]       # the default schedule runs 100 critic steps while it < 10
    torch.cuda.synchronize()
    mem.check()
    assert int(mem.ctl[0]) == calls0 + 4 * (2 + 1 + t._it["citers"])      # draw + replace + draws + 5 critic draws per iteration
    assert t.counter_c >= 4 * t._it["citers"] and t.counter_g >= 4
    assert bool(torch.isfinite(t.gv.flat).all()) and bool(torch.isfinite(t.cri.flat).all())
    assert float((t.gv.flat - p0).abs().max()) > 0
    assert all(bool(torch.isfinite(o["g_loss"])) for o in outs)
    return t.gv.flat.clone(), t.cri.flat.clone(), t.graph_launches["iteration"]

  a = run()
  b = run()
  assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]), "the graph-replayed iterations are not reproducible"
  assert a[2] > 100
