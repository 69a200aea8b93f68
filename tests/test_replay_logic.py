"""CPU: the device replay memory's selection logic (exposure_b200/csrc/replay_logic.cuh, compiled for the host by
tests/host_math/replay_harness.cpp) obeys the semantics of replay_memory.py:187-273 and behaves like the
reference-faithful host ReplayMemory (which tests/test_replay_reference.py pins draw for draw to the reference's own
code) over hundreds of iterations: same pool composition statistics, same batch invariants."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
S = 11
STOPPED, STEP = 1, 2


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


class LogicPool:
  """The record buffer of the device replay memory, on numpy: pool | generator outputs | fresh."""

  def __init__(self, rl, P, B, test_steps, max_len, keep, seed):
    self.rl, self.P, self.B, self.test_steps, self.max_len, self.keep, self.seed = rl, P, B, test_steps, max_len, keep, seed
    self.call = 0
    self.next_id = 0
    self.ids = np.zeros(2 * P + 2 * B, np.int64)            # record identity (stands for the image)
    self.states = np.zeros((2 * P + 2 * B, S), np.float32)
    self._fresh()
    self.ids[:P] = self.ids[P + B:P + B + P]                # initial fill_pool
    self._fresh()

  def _fresh(self):
    F = self.P + self.B
    self.ids[self.P + self.B:] = np.arange(self.next_id, self.next_id + F)
    self.next_id += F
    self.states[self.P + self.B:] = 0

  def generator_batch(self):
    P, B = self.P, self.B
    batch = np.zeros(B, np.int64); rest = np.zeros(P, np.int32); fu = ctypes.c_int()
    pool_states = np.ascontiguousarray(self.states[:P])
    n_rest = self.rl.rl_draw_generator(_p(pool_states), S, P, B, ctypes.c_ulonglong(self.seed), ctypes.c_ulonglong(self.call),
                                       _p(batch), _p(rest), ctypes.byref(fu))
    self.call += 1
    self._hand = (rest, n_rest, fu.value)
    return batch

  def replace(self, new_states):
    P, B = self.P, self.B
    rest, n_rest, fu = self._hand
    self.states[P:P + B] = new_states
    src = np.zeros(P, np.int64)
    ns = np.ascontiguousarray(new_states, dtype=np.float32)
    self.rl.rl_replace(_p(ns), S, P, B, self.max_len, ctypes.c_float(self.keep), ctypes.c_ulonglong(self.seed),
                       ctypes.c_ulonglong(self.call), _p(rest), n_rest, fu, _p(src))
    self.call += 1
    self.ids[:P], self.states[:P] = self.ids[src], self.states[src]
    self._fresh()
    return src

  def critic_batch(self):
    batch = np.zeros(self.B, np.int64)
    pool_states = np.ascontiguousarray(self.states[:self.P])
    nt = self.rl.rl_draw_critic(_p(pool_states), S, self.P, self.B, ctypes.c_ulonglong(self.seed), ctypes.c_ulonglong(self.call),
                                _p(batch))
    self.call += 1
    return batch, nt


def _advance(states, test_steps):
  """agent.py:208-222: the state update of one generator step."""
  new = states.copy()
  last = (np.abs(states[:, STEP] + 1 - test_steps) < 1e-4).astype(np.float32)
  new[:, 0] = last
  new[:, STOPPED] = last
  new[:, STEP] = states[:, STEP] + 1
  return new


def test_semantics_of_each_operation(rl):
  P, B = 128, 64
  m = LogicPool(rl, P, B, test_steps=5, max_len=7, keep=0.5, seed=11)
  seen_ids = set(m.ids[:P].tolist())
  for it in range(300):
    pool_before = m.ids[:P].copy()
    stopped_before = m.states[:P, STOPPED] > 0
    batch = m.generator_batch()
    rest, n_rest, fu = m._hand
    # the batch: B distinct records, none terminated; batch + rest + dropped = the pool (when it did not run dry)
    assert len(set(batch.tolist())) == B
    assert not (m.states[batch, STOPPED] > 0).any()
    if fu == 0:
      in_pool = batch[batch < P]
      assert len(in_pool) == B
      dropped = set(range(P)) - set(in_pool.tolist()) - set(rest[:n_rest].tolist())
      assert all(stopped_before[d] for d in dropped), "a non-terminated record was dropped"
      assert len(set(in_pool.tolist()) & set(rest[:n_rest].tolist())) == 0
    else:
      assert fu == P and n_rest == P - (batch >= P + B).sum() and (rest[:n_rest] >= P + B).all()
    new_states = _advance(m.states[batch], m.test_steps)
    out_ids = m.ids[batch].copy()
    m.ids[P:P + B] = out_ids
    src = m.replace(new_states)
    # the new pool: P distinct records; every remaining record is there; outputs with step < max_len are all there
    assert len(set(src.tolist())) == P
    assert set(rest[:min(n_rest, P)].tolist()) <= set(src.tolist())
    young = {P + j for j in range(B) if new_states[j, STEP] < m.max_len}
    if n_rest + len(young) <= P:
      assert young <= set(src.tolist())
    assert len(set(m.ids[:P].tolist())) == P
    cb, nt = m.critic_batch()
    if nt > 0:
      assert (m.states[cb, STOPPED] > 0).all() and (cb < P).all()
      assert len(set(cb.tolist())) == min(nt, B), "cycles through the terminated records in order"
  assert (m.states[:P, STOPPED] > 0).sum() > 0


def test_pool_statistics_match_the_reference_faithful_host_memory(rl):
  """Mean trajectory step and terminated fraction of the pool in the stationary regime: device logic vs the host
  ReplayMemory (draw-for-draw equal to the reference) driven by the same state update."""
  from exposure_b200.replay import ReplayMemory
  from exposure_b200.trainer import default_cfg
  from exposure_b200.util import STATE_STEP_DIM, STATE_STOPPED_DIM
  P, B = 128, 64
  cfg = default_cfg()
  cfg.source_img_size = cfg.real_img_size = 4

  class Tiny:
    def get_next_batch(self, n):
      return torch.zeros(n, 4, 4, 3)

  host = ReplayMemory(cfg, Tiny(), Tiny(), torch.device("cpu"), seed=5)
  dev = LogicPool(rl, P, B, test_steps=cfg.test_steps, max_len=cfg.maximum_trajectory_length, keep=cfg.over_length_keep_prob, seed=5)
  hs, ht, ds, dt = [], [], [], []
  for it in range(600):
    img, st, slots = host.get_next_fake_batch(B)
    host.replace_memory(img, torch.from_numpy(_advance(st.numpy(), cfg.test_steps)), slots)
    batch = dev.generator_batch()
    dev.replace(_advance(dev.states[batch], cfg.test_steps))
    if it >= 100:
      hs.append(np.mean([host.step[s] for s in host.image_pool])); ht.append(np.mean([host.stopped[s] for s in host.image_pool]))
      ds.append(dev.states[:P, STEP].mean()); dt.append((dev.states[:P, STOPPED] > 0).mean())
  assert abs(np.mean(hs) - np.mean(ds)) < 0.05 * np.mean(hs), (np.mean(hs), np.mean(ds))
  assert abs(np.mean(ht) - np.mean(dt)) < 0.1 * np.mean(ht) + 0.005, (np.mean(ht), np.mean(dt))


def test_philox_uniforms_are_uniform_and_reproducible(rl):
  a, b = np.zeros(20000, np.float32), np.zeros(20000, np.float32)
  rl.rl_uniforms(ctypes.c_ulonglong(3), ctypes.c_ulonglong(7), _p(a), 20000)
  rl.rl_uniforms(ctypes.c_ulonglong(3), ctypes.c_ulonglong(7), _p(b), 20000)
  assert (a == b).all() and a.min() >= 0 and a.max() < 1
  assert abs(a.mean() - 0.5) < 0.01 and abs(a.var() - 1 / 12) < 0.005
  rl.rl_uniforms(ctypes.c_ulonglong(3), ctypes.c_ulonglong(8), _p(b), 20000)
  assert (a != b).mean() > 0.99


def test_resident_provider_cycles_pregenerated_batches():
  """ResidentProvider: `slots` batches per size generated once, then handed out round robin (bench.py's data side)."""
  import torch
  from exposure_b200.replay import ResidentProvider

  class Counting:
    def __init__(self):
      self.calls = 0
    def get_next_batch(self, n):
      self.calls += 1
      return torch.full((n, 2), float(self.calls))

  inner = Counting()
  p = ResidentProvider(inner, slots=3)
  p.prefill(4)
  assert inner.calls == 3
  got = [float(p.get_next_batch(4)[0, 0]) for _ in range(7)]
  assert got == [1.0, 2.0, 3.0, 1.0, 2.0, 3.0, 1.0] and inner.calls == 3
  first = [float(p.get_next_batch(5)[0, 0]) for _ in range(4)]          # another size: its own ring, filled lazily
  assert first == [4.0, 5.0, 6.0, 5.0] or first == [4.0, 5.0, 6.0, 4.0]
