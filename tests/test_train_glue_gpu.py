"""GPU: the fused bookkeeping kernels of the train step (csrc/train_glue.cu) against plain torch expressions of
the reference lines they replace (net.py:164-187, agent.py:113-125, filters.py:39-44, critics.py:48-87)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K(built_lib):
  assert torch.cuda.is_available()
  from exposure_b200 import nn_ops
  return nn_ops


def test_colsum_multi_matches_sum(K):
  g = torch.Generator(device="cuda").manual_seed(0)
  shapes = [(8 * 1024, 32), (2048, 64), (512, 128), (128, 256), (8, 1024), (8, 1), (300, 7), (16, 128)]
  src = [torch.randn(r, c, device="cuda", generator=g) for r, c in shapes]
  dst = [torch.full((c,), float("nan"), device="cuda") for _, c in shapes]
  K.colsum_multi([(s, d, False) for s, d in zip(src, dst)])
  torch.cuda.synchronize()
  for s, d in zip(src, dst):
    ref = s.double().sum(dim=0)
    assert torch.allclose(d.double(), ref, rtol=1e-5, atol=1e-4 * float(ref.abs().max()) + 1e-6), tuple(s.shape)
  # a task list with FEWER column blocks first, then one with more, on the same workspace (the ticket block is fixed)
  small = torch.randn(64, 32, device="cuda", generator=g)
  big = torch.randn(8, 1024, device="cuda", generator=g)
  o_s, o_b = torch.empty(32, device="cuda"), torch.empty(1024, device="cuda")
  K.colsum_multi([(small, o_s, False)])
  K.colsum_multi([(big, o_b, False)])
  assert torch.allclose(o_b, big.sum(dim=0), rtol=1e-5, atol=1e-5) and torch.allclose(o_s, small.sum(dim=0), rtol=1e-5, atol=1e-5)
  # accumulate, 4-D sources, more than 8 tasks (two launches), repeated use of the self-cleaning workspace
  src4 = [torch.randn(3, 5, 7, 16, device="cuda", generator=g) for _ in range(10)]
  dst4 = [torch.ones(16, device="cuda") for _ in range(10)]
  for _ in range(2):
    K.colsum_multi([(s, d, True) for s, d in zip(src4, dst4)])
  torch.cuda.synchronize()
  for s, d in zip(src4, dst4):
    assert torch.allclose(d, 1 + 2 * s.reshape(-1, 16).sum(dim=0), rtol=1e-5, atol=1e-5)
  # deterministic
  a = torch.randn(5000, 96, device="cuda", generator=g)
  o1, o2 = torch.empty(96, device="cuda"), torch.empty(96, device="cuda")
  K.colsum_multi([(a, o1, False)])
  K.colsum_multi([(a, o2, False)])
  assert torch.equal(o1, o2)


def test_critic_inputs_and_scalars(K):
  g = torch.Generator(device="cuda").manual_seed(1)
  B = 6
  real, fake = torch.rand(B, 64, 64, 3, device="cuda", generator=g), torch.rand(B, 64, 64, 3, device="cuda", generator=g)
  alpha = torch.rand(B, device="cuda", generator=g)
  X = K.critic_inputs(real, fake, alpha)
  assert torch.equal(X[:B], real) and torch.equal(X[B:2 * B], fake)
  ref = real + alpha[:, None, None, None] * (fake - real)                       # net.py:177-179
  assert torch.allclose(X[2 * B:], ref, rtol=0, atol=1e-7)
  logits = torch.randn(3 * B, device="cuda", generator=g) * 5
  norm = torch.rand(B, device="cuda", generator=g) * 3
  ema = torch.tensor([0.3, 0.1, 4.0], device="cuda")
  e0 = ema.clone()
  out = K.critic_scalars(logits, norm, 10.0, ema, 0.99)
  torch.cuda.synchronize()
  emd = logits[:B].mean() - logits[B:2 * B].mean()                              # net.py:164
  gp = 10.0 * (torch.clamp(norm - 1.0, min=0.0) ** 2).mean()                    # net.py:185-187
  cav = (logits[:B].mean() + logits[B:2 * B].mean()) * 0.5                      # net.py:165
  want = torch.stack([emd, gp, norm.mean(), -emd + gp, cav])
  assert torch.allclose(out[:5], want, rtol=1e-5, atol=1e-6)
  biased = e0[1] * 0.99 + cav * 0.01                                            # moving_averages._zero_debias
  assert torch.allclose(ema, torch.stack([biased / (1 - 0.99 ** 5.0), biased, torch.tensor(5.0, device="cuda")]), rtol=1e-5)
  before = ema.clone()
  K.critic_scalars(logits, norm, 10.0, None)                                    # apply=False: the average does not move
  assert torch.equal(ema, before)


def _heads(K, seed=3):
  from exposure_b200.nets import FC1, MASK_PARAMS, NUM_PARAMS, ParamStore, PolicyNet
  store = ParamStore(torch.device("cuda"))
  net = PolicyNet(store, n_states=11, scope="generator")
  store.finalize(seed)
  g = torch.Generator(device="cuda").manual_seed(seed)
  for name in net.fc2:                                                          # biases start at zero: make them count
    store.p[name + "/biases"].copy_(torch.randn(store.p[name + "/biases"].shape, device="cuda", generator=g) * 0.1)
  return store, net, FC1, MASK_PARAMS, list(NUM_PARAMS), g


@pytest.mark.parametrize("masking", [False, True])
def test_heads_forward_select_backward(K, masking):
  store, net, FC1, NM, NP, g = _heads(K)
  B = 9
  H = torch.randn(B, 8 * FC1, device="cuda", generator=g)
  ids = torch.tensor([0, 7, 4, -1, 2, 7, 3, 5, 1], dtype=torch.int32, device="cuda")
  L = net.heads()
  O = K.heads_fc2_fwd(L, H)
  sel, msel = K.heads_select(L, O, ids, 24, masking)
  gsel = torch.randn(B, 24, device="cuda", generator=g)
  gmsel = torch.randn(B, NM, device="cuda", generator=g) if masking else None
  for b in range(B):                                                            # entries >= n carry no gradient (regressor contract)
    if int(ids[b]) >= 0:
      gsel[b, NP[int(ids[b])]:] = 0
  store.grad.fill_(float("nan"))
  dH = K.heads_fc2_bwd(L, H, ids, gsel, gmsel)
  torch.cuda.synchronize()
  # reference: filters.py:39-44 per head, one-hot select agent.py:113-125, autograd
  Hr = H.double().requires_grad_(True)
  Ws = [store.p[n + "/weights"].double().requires_grad_(True) for n in net.fc2]
  bs = [store.p[n + "/biases"].double().requires_grad_(True) for n in net.fc2]
  loss = 0
  for b in range(B):
    j = int(ids[b])
    o_all = [Hr[b, k * FC1:(k + 1) * FC1] @ Ws[k] + bs[k] for k in range(8)]
    for k in range(8):
      assert torch.allclose(O[b, k, :o_all[k].numel()].double(), o_all[k].detach(), rtol=1e-5, atol=1e-6)
      assert o_all[k].numel() == O.shape[2] or float(O[b, k, o_all[k].numel():].abs().max()) == 0
    if j < 0:
      assert float(sel[b].abs().max()) == 0
      continue
    n = NP[j]
    assert torch.allclose(sel[b, :n].double(), o_all[j][:n].detach(), rtol=1e-5, atol=1e-6)
    assert n == 24 or float(sel[b, n:].abs().max()) == 0
    loss = loss + (o_all[j][:n] * gsel[b, :n].double()).sum()
    if masking:
      assert torch.allclose(msel[b].double(), o_all[j][n:n + NM].detach(), rtol=1e-5, atol=1e-6)
      loss = loss + (o_all[j][n:n + NM] * gmsel[b].double()).sum()
  loss.backward()
  # dH carries lrelu'(H) (the fc1 output is stored post-activation)
  dl = torch.where(H > 0, 1.0, torch.where(H < 0, 0.2, 0.6)).double()
  assert torch.allclose(dH.double(), Hr.grad * dl, rtol=1e-5, atol=1e-6)
  for k, name in enumerate(net.fc2):
    gw = store.g[name + "/weights"].double()
    gb = store.g[name + "/biases"].double()
    rw = Ws[k].grad if Ws[k].grad is not None else torch.zeros_like(Ws[k])
    rb = bs[k].grad if bs[k].grad is not None else torch.zeros_like(bs[k])
    assert torch.allclose(gw, rw, rtol=1e-5, atol=1e-6), name
    assert torch.allclose(gb, rb, rtol=1e-5, atol=1e-6), name


def test_stats_bwd_gin_equals_the_composed_launches(K):
  g = torch.Generator(device="cuda").manual_seed(2)
  for cin in (6, 17):
    B = 5
    img = torch.rand(B, 64, 64, 3, device="cuda", generator=g) * 1.3 - 0.1
    stats = K.stats_fwd(img)
    g_in = torch.randn(B, 64, 64, cin, device="cuda", generator=g)
    got = K.stats_bwd_gin(img, stats, g_in)
    g_vec = K.colsum(g_in, batch=B).reshape(B, -1)
    ref = K.stats_bwd(img, stats, g_vec[:, -3:].contiguous(), g_direct=g_in[..., :3].contiguous())
    torch.cuda.synchronize()
    assert torch.allclose(got, ref, rtol=1e-5, atol=1e-6 * float(ref.abs().max()))


def test_set_floats(K):
  """exp_set_floats: up to 16 scalars by value in one launch (the per-iteration lr_t / progress of the captured iteration)."""
  dst = torch.full((20,), -1.0, device="cuda")
  vals = [0.125 * i - 0.5 for i in range(16)]
  K.set_floats(dst, vals)
  torch.cuda.synchronize()
  assert dst[:16].cpu().tolist() == vals and dst[16:].cpu().tolist() == [-1.0] * 4
  K.set_floats(dst[3:], [7.0])
  assert float(dst[3]) == 7.0 and float(dst[4]) == vals[4]
  with pytest.raises(Exception):
    K.set_floats(dst, [0.0] * 17)

