#!/usr/bin/env python
"""Host<->device copy bandwidth of this box for the e2e leg's buffer size: H2D alone, D2H alone and
both directions at once (pinned memory, two streams).  The concurrent figure is the ceiling of
bench.py's chain8 `e2e` number."""
import json
import sys
import torch

mb = int(sys.argv[1]) if len(sys.argv) > 1 else 201
n = mb * 1000 * 1000 // 4
h_in = torch.empty(n, dtype=torch.float32).pin_memory()
h_out = torch.empty(n, dtype=torch.float32).pin_memory()
d_in = torch.empty(n, device="cuda")
d_out = torch.ones(n, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=10):
  fn()
  torch.cuda.synchronize()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(reps):
    fn()
  for s in (s1, s2):
    torch.cuda.current_stream().wait_stream(s)
  b.record()
  torch.cuda.synchronize()
  return a.elapsed_time(b) / reps


def h2d():
  with torch.cuda.stream(s1):
    d_in.copy_(h_in, non_blocking=True)


def d2h():
  with torch.cuda.stream(s2):
    h_out.copy_(d_out, non_blocking=True)


def both():
  h2d()
  d2h()


res = {"mb": mb}
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
  ms = timed(fn)
  res[name + "_ms"] = ms
  res[name + "_gbs_per_direction"] = n * 4 / ms / 1e6
print(json.dumps(res))
