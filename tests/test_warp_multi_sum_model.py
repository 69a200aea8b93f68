"""CPU model of the multi-value warp reduction used by the fused chain kernel
(exposure_b200/csrc/filters.cu warp_multi_sum): N per-lane values are summed over the 32 lanes in
N-1 + max(0, 5 - log2 N) shuffle stages' worth of exchanges instead of 5 N.  The model executes the same
stage schedule on 32 simulated lanes and checks (a) every lane ends with the warp total of the value index
it reports, (b) the lanes the kernel lets write -- (lane & (32/N - 1)) == 0 -- cover every index exactly once,
(c) the exchange count."""
import numpy as np
import pytest


def warp_multi_sum_model(vals):
  """vals [32 lanes, N] -> (result per lane, idx per lane, number of shuffles issued per lane)."""
  lanes, N = vals.shape
  v = [list(map(float, vals[l])) for l in range(lanes)]
  idx = [0] * lanes
  n = N
  shuffles = 0
  d = 16
  while d >= 1:
    if n > 1:
      h = n // 2
      new = [row[:] for row in v]
      for l in range(lanes):
        up = (l & d) != 0
        p = l ^ d
        p_up = (p & d) != 0
        for i in range(h):
          keep = v[l][i + h] if up else v[l][i]
          recv = v[p][i] if p_up else v[p][i + h]          # what the partner sends
          new[l][i] = keep + recv
        if up:
          idx[l] += h
      v = new
      shuffles += h
      n = h
    else:
      v = [[v[l][0] + v[l ^ d][0]] + v[l][1:] for l in range(lanes)]
      shuffles += 1
    d //= 2
  return [row[0] for row in v], idx, shuffles


@pytest.mark.parametrize("N", [1, 2, 4, 8, 16, 32])
def test_model(N):
  rng = np.random.RandomState(N)
  vals = rng.randint(-50, 50, size=(32, N)).astype(np.float64)      # integers: exact sums, order-independent
  res, idx, shuffles = warp_multi_sum_model(vals)
  totals = vals.sum(axis=0)
  for l in range(32):
    assert res[l] == totals[idx[l]], (l, idx[l])
  writers = [l for l in range(32) if (l & (32 // N - 1)) == 0]
  assert sorted(idx[l] for l in writers) == list(range(N))
  log2n = int(np.log2(N))
  assert shuffles == (N - 1) + (5 - log2n)                          # e.g. 8 values: 4+2+1 halving + 2 plain = 9, not 40
