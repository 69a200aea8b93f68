"""Replay memory with the image pool resident in HBM.

Semantics of replay_memory.py:8-282 (pool of cfg.replay_memory_size records; generator batches
pop non-terminated records and drop terminated ones; critic batches read terminated records;
outputs are re-inserted while step < maximum_trajectory_length, else with probability
over_length_keep_prob; the pool is topped up with fresh RAW images), but records are SLOTS of two
device tensors -- images [cap,64,64,3] and states [cap,11] -- and a step costs one index upload
plus gather/scatter on the device instead of O(pool) Python list slicing, np.stack and an
H2D + D2H of every image.  The host keeps only what it can know without reading the device:
`step` and `stopped` evolve deterministically (agent.py:208-222), so the reference's selection
logic runs on host integers with no synchronisation."""
import random

import torch

from .util import STATE_STEP_DIM, STATE_STOPPED_DIM


class SyntheticProvider:
  """Stand-in for FiveKDataProvider / ArtistDataProvider (out of scope): seeded
  linear-RGB-like 64x64 batches generated on the device (SURVEY 8d statistics)."""

  def __init__(self, device, kind="raw", seed=0, size=64):
    self.device, self.kind, self.size = device, kind, size
    self.gen = torch.Generator(device=device).manual_seed(seed)

  def get_next_batch(self, n):
    s = self.size
    z = torch.randn(n, s, s, 3, device=self.device, generator=self.gen)
    if self.kind == "raw":       # dark linear RAW: median ~0.04
      return torch.exp(z - 3.2).clamp_(0, 4)
    return torch.exp(0.7 * z - 1.2).clamp_(0, 1)   # retouched targets: brighter, display range


class ResidentProvider:
  """`slots` batches of `inner` per batch size, generated once and kept in HBM, then handed out round robin: the
  data-set side of a measurement whose inputs are resident when the timed region starts (bench.py; a provider that
  draws new random batches costs ~25 small kernels per train iteration that are not part of the path)."""

  def __init__(self, inner, slots=16):
    self.inner, self.slots = inner, int(slots)
    self._rings, self._next = {}, {}

  def prefill(self, n):
    ring = self._rings.setdefault(n, [])
    while len(ring) < self.slots:
      ring.append(self.inner.get_next_batch(n))

  def get_next_batch(self, n):
    ring = self._rings.setdefault(n, [])
    k = self._next.get(n, 0)
    self._next[n] = k + 1
    if len(ring) < self.slots:
      ring.append(self.inner.get_next_batch(n))
      return ring[-1]
    return ring[k % self.slots]


class DeviceReplayMemory:
  """The replay memory with its SELECTION LOGIC on the device too (csrc/replay.cu): every operation
  of replay_memory.py:187-273 is an index list computed by a one-thread kernel over the pool's 128 state rows plus a
  gather, so a train iteration needs no host list handling and no index upload and can be captured as ONE CUDA graph
  (Trainer.enable_iteration_graph).  Random choices come from Philox streams keyed by `seed` and a device-side call
  counter: reproducible for a seed, but not the reference's Mersenne-Twister sequence -- ReplayMemory below keeps
  reproducing THAT draw for draw (host mode, tests/test_replay_reference.py); tests/test_replay_logic.py shows both
  produce the same pool statistics.

  Records live in one buffer: [0, P) pool | [P, P+B) outputs of the last generator step | [P+B, 2P+2B) fresh RAW
  records of this iteration.  Same method names as the reference class (get_next_fake_batch / replace_memory /
  replay_fake_batch); the graph path uses stage_fresh / draw_generator / replace / draw_critic directly."""

  def __init__(self, cfg, fake_provider, real_provider, device, seed=0):
    from . import _cabi
    self.cfg = cfg
    self.fake_dataset, self.real_dataset = fake_provider, real_provider
    self.device = device
    self.seed = int(seed) * 2654435761 + 12345
    self.P, self.B, self.S = int(cfg.replay_memory_size), int(cfg.batch_size), int(cfg.num_state_dim)
    self.target_pool_size = self.P
    P, B = self.P, self.B
    self.F = P + B
    s = cfg.source_img_size
    total = P + B + self.F
    self.images = torch.zeros(total, s, s, 3, device=device)
    self.states = torch.zeros(total, self.S, device=device)       # the fresh region's states stay zero (get_initial_states)
    self.ctl = torch.zeros(_cabi.lib().exp_replay_ctl_words(), dtype=torch.int32, device=device)
    self.rest_src = torch.zeros(P, dtype=torch.int32, device=device)
    self.batch_src = torch.zeros(B, dtype=torch.int64, device=device)
    self.crit_src = torch.zeros(B, dtype=torch.int64, device=device)
    self.pool_src = torch.zeros(P, dtype=torch.int64, device=device)
    self._tmp_images = torch.empty(P, s, s, 3, device=device)
    self._tmp_states = torch.empty(P, self.S, device=device)
    self._gen_images = torch.empty(B, s, s, 3, device=device)
    self._gen_states = torch.empty(B, self.S, device=device)
    self._crit_images = torch.empty(B, s, s, 3, device=device)
    self.stage_fresh()                                            # initial fill_pool (replay_memory.py:42-43)
    self.images[:P].copy_(self.images[P + B:P + B + P])
    self.states[:P].zero_()

  # ---- device operations (no host synchronisation; all capturable) ------------------------------------------
  def stage_fresh(self):
    """Fresh RAW records of this iteration from the data provider (the only host-facing step; NOT inside a graph)."""
    self.images[self.P + self.B:].copy_(self.fake_dataset.get_next_batch(self.F))

  def draw_generator(self):
    from . import nn_ops as K
    K.replay_draw_generator(self.states[:self.P], self.B, self.seed, self.ctl, self.batch_src, self.rest_src)
    K.gather_rows(self.images, self.batch_src, self._gen_images)
    K.gather_rows(self.states, self.batch_src, self._gen_states)
    return self._gen_images, self._gen_states

  def replace(self, new_images, new_states):
    from . import nn_ops as K
    P, B = self.P, self.B
    self.images[P:P + B].copy_(new_images)
    self.states[P:P + B].copy_(new_states)
    K.replay_replace(self.states[P:P + B], P, self.cfg.maximum_trajectory_length, self.cfg.over_length_keep_prob, self.seed,
                     self.ctl, self.rest_src, self.pool_src)
    K.gather_rows(self.images, self.pool_src, self._tmp_images)
    K.gather_rows(self.states, self.pool_src, self._tmp_states)
    self.images[:P].copy_(self._tmp_images)
    self.states[:P].copy_(self._tmp_states)

  def draw_critic(self):
    from . import nn_ops as K
    K.replay_draw_critic(self.states[:self.P], self.B, self.seed, self.ctl, self.crit_src)
    return K.gather_rows(self.images, self.crit_src, self._crit_images)

  def check(self):
    """Host synchronisation: raises like the reference's assertion if a critic batch found no terminated record."""
    if int(self.ctl[4]) != 0:
      raise AssertionError("No terminated states discovered")          # replay_memory.py:258-259

  # ---- the reference's method names (eager use: GAN.train, Trainer.train_iteration) ------------------------------
  def get_next_fake_batch(self, batch_size):
    assert batch_size == self.B
    self.stage_fresh()
    img, st = self.draw_generator()
    return img, st, None

  def replace_memory(self, new_images, new_states, old_slots=None):
    self.replace(new_images, new_states)

  def replay_fake_batch(self, batch_size):
    assert batch_size == self.B
    return self.draw_critic(), None

  def fill_pool(self):
    pass                                   # the pool is always full: replace() tops it up (replay_memory.py:196)

  def debug(self):
    st = self.states[:self.P]
    return self.P, float(st[:, STATE_STEP_DIM].mean())


class ReplayMemory:

  def __init__(self, cfg, fake_provider, real_provider, device, seed=0):
    self.cfg = cfg
    self.fake_dataset = fake_provider
    self.real_dataset = real_provider
    self.device = device
    self.rng = random.Random(seed)
    self.target_pool_size = cfg.replay_memory_size
    cap = cfg.replay_memory_size + 2 * cfg.batch_size
    s = cfg.source_img_size
    self.images = torch.zeros(cap, s, s, 3, device=device)
    self.states = torch.zeros(cap, cfg.num_state_dim, device=device)
    self.free = list(range(cap))
    self.image_pool = []                 # slot ids, in the reference's list order
    self.step = [0] * cap                # host mirror of states[:, STATE_STEP_DIM]
    self.stopped = [False] * cap         # host mirror of states[:, STATE_STOPPED_DIM]
    self._outstanding = []               # slots handed out by get_next_fake_batch and not yet replaced
    self.fill_pool()

  def _idx(self, slots):
    return torch.tensor(slots, dtype=torch.long).to(self.device, non_blocking=True)

  def _release(self, slots):
    self.free.extend(slots)

  def fill_pool(self):                   # replay_memory.py:64-75
    while len(self.image_pool) < self.target_pool_size:
      batch = self.fake_dataset.get_next_batch(self.cfg.batch_size)
      slots = [self.free.pop() for _ in range(batch.shape[0])]
      idx = self._idx(slots)
      self.images.index_copy_(0, idx, batch)
      self.states.index_fill_(0, idx, 0.0)
      for sl in slots:
        self.step[sl], self.stopped[sl] = 0, False
      self.image_pool.extend(slots)
    self._release(self.image_pool[self.target_pool_size:])
    self.image_pool = self.image_pool[:self.target_pool_size]

  def get_next_fake_batch(self, batch_size):     # replay_memory.py:230-246
    # the reference's records simply leave the pool here; a caller that never hands them back through
    # replace_memory (an eval / bench loop) must not leak their slots
    if self._outstanding:
      self._release(self._outstanding)
      self._outstanding = []
    self.rng.shuffle(self.image_pool)
    batch = []
    while len(batch) < batch_size:
      if not self.image_pool:
        self.fill_pool()
      slot = self.image_pool.pop(0)
      if not self.stopped[slot]:
        batch.append(slot)
      else:
        self._release([slot])            # finished images are dropped here
    idx = self._idx(batch)
    self._outstanding = list(batch)
    return self.images.index_select(0, idx), self.states.index_select(0, idx), batch

  def replace_memory(self, new_images, new_states, old_slots):    # replay_memory.py:187-196
    """new_images / new_states: device outputs of the generator step for the records `old_slots`."""
    if sorted(old_slots) != sorted(self._outstanding):
      raise ValueError("replace_memory() takes the slots of the most recent get_next_fake_batch()")
    self._outstanding = []
    self.rng.shuffle(self.image_pool)
    keep_rows, keep_slots = [], []
    for row, slot in enumerate(old_slots):
      step = self.step[slot] + 1
      stopped = abs(step - self.cfg.test_steps) < 1e-4          # agent.py:210-218
      if step < self.cfg.maximum_trajectory_length or self.rng.random() < self.cfg.over_length_keep_prob:
        self.step[slot], self.stopped[slot] = step, stopped
        keep_rows.append(row)
        keep_slots.append(slot)
      else:
        self._release([slot])
    if keep_slots:
      rows, idx = self._idx(keep_rows), self._idx(keep_slots)
      self.images.index_copy_(0, idx, new_images.index_select(0, rows))
      self.states.index_copy_(0, idx, new_states.index_select(0, rows))
      self.image_pool.extend(keep_slots)
    self.fill_pool()
    self.rng.shuffle(self.image_pool)

  def replay_fake_batch(self, batch_size):       # replay_memory.py:249-273
    self.fill_pool()
    self.rng.shuffle(self.image_pool)
    batch = []
    counter = 0
    while len(batch) < batch_size:
      counter += 1
      assert counter <= batch_size * 10, "No terminated states discovered"
      for slot in self.image_pool:
        if self.stopped[slot]:
          batch.append(slot)
          if len(batch) >= batch_size:
            break
    idx = self._idx(batch)
    return self.images.index_select(0, idx), self.states.index_select(0, idx)

  def debug(self):                               # replay_memory.py:275-282
    tot = sum(self.step[s] for s in self.image_pool)
    return len(self.image_pool), 1.0 * tot / max(1, len(self.image_pool))
