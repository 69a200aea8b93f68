"""N-step filter chain driver (BASELINE.json configs[1]/[4]: "8-filter chain fwd+bwd").

The chain x_0 -> f_{id_0} -> x_1 -> ... -> x_N is a benchmark construct built from the
reference's per-filter maths (filters.py process()); in the reference one step applies one
selected filter (agent.py:113-125) and an episode is cfg.test_steps such steps.  Each step is
ONE fused CUDA kernel forward and ONE backward (exposure_b200/csrc/filters*.cu): the
filter_param_regressor runs in the kernel prologue (EXP_OPT_LOGITS) and its chain rule in the
backward's finishing CTA, so a chain step is exactly 2N launches.  Activations x_0..x_{N-1} stay
resident in HBM for the backward (12 B/pixel/step); outputs are recomputed, never re-read.

Explicit schedule, no autograd: forward() then backward(); capture() records both into one CUDA
graph over the chain's static buffers."""
import torch

from . import ops


class FilterChain:

  def __init__(self, ids, variant=ops.VARIANT_AUTO, fused_regressor=True):
    """ids: list of N entries, each an int filter id (uniform over the batch) or a CUDA int32
    tensor [B] of per-image ids."""
    self.ids = list(ids)
    self.variant = variant
    self.fused = fused_regressor
    self._acts = None
    self._params = None
    self._logits = None
    self._gbuf = None
    self._glog = None
    self._graph = None

  def _alloc(self, shape, device):
    n = len(self.ids)
    if self._acts is None or tuple(self._acts[0].shape) != tuple(shape) or self._acts[0].device != device:
      mk = lambda: torch.empty(shape, device=device, dtype=torch.float32)
      self._acts = [mk() for _ in range(n + 1)]
      self._gbuf = [mk() for _ in range(2)]
      self._glog = None
      self._graph = None

  def input_buffer(self, shape, device):
    self._alloc(shape, device)
    return self._acts[0]

  def _nparams(self, k):
    fid = self.ids[k]
    return ops.NUM_PARAMS_ALL[fid] if isinstance(fid, int) else ops.PSTRIDE

  def forward(self, x, logits_list):
    """x: [B,H,W,3]; logits_list[k]: [B, >= n_k] raw regressor inputs.  Returns x_N."""
    assert len(logits_list) == len(self.ids)
    self._alloc(x.shape, x.device)
    if self._acts[0].data_ptr() != x.data_ptr():
      self._acts[0].copy_(x)
    self._logits = [l.contiguous() for l in logits_list]
    if self._glog is None or any(g.shape != l.shape for g, l in zip(self._glog, self._logits)):
      self._glog = [torch.zeros_like(l) if self.fused else torch.zeros(l.shape[0], ops.PSTRIDE, device=l.device)
                    for l in self._logits]
    self._params = []
    for k, fid in enumerate(self.ids):
      if self.fused:
        p = self._logits[k]
      else:
        p = ops.filter_regress_fwd(self._logits[k], fid)
      self._params.append(p)
      ops.filter_fwd(self._acts[k], p, fid, out=self._acts[k + 1], variant=self.variant, logits=self.fused)
    return self._acts[-1]

  def forward_resident(self, logits_list):
    """forward() with x_0 already resident in input_buffer()."""
    return self.forward(self._acts[0], logits_list)

  def backward(self, gout, need_input_grad=True):
    """gout = dL/dx_N.  Returns (dL/dx_0 or None, [dL/dlogits_k])."""
    n = len(self.ids)
    g = gout
    glogits = [None] * n
    for k in reversed(range(n)):
      fid = self.ids[k]
      need_gx = need_input_grad or k > 0
      gx, gp = ops.filter_bwd(self._acts[k], g, self._params[k], fid, need_gx=need_gx,
                              gx_out=self._gbuf[k & 1] if need_gx else None, variant=self.variant,
                              logits=self.fused, gparams_out=self._glog[k])
      if self.fused:
        glogits[k] = gp
      else:
        glogits[k] = ops.filter_regress_bwd(self._logits[k], gp, fid)
      g = gx
    return g, glogits

  # ---- CUDA graph of one fwd+bwd over the static buffers --------------------------------
  def capture(self, logits_list, gout):
    """Record forward_resident(logits_list) + backward(gout) into a CUDA graph.  `logits_list`
    and `gout` become the graph's static inputs (update them in place), x_0 is input_buffer().
    Returns (y, gx, glogits) static outputs."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(2):
        self.forward_resident(logits_list)
        self.backward(gout)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    self._graph = torch.cuda.CUDAGraph()
    l0 = ops.launch_count
    with torch.cuda.graph(self._graph):
      y = self.forward_resident(logits_list)
      gx, gl = self.backward(gout)
    self.graph_launches = ops.launch_count - l0
    self._graph_out = (y, gx, gl)
    return self._graph_out

  def replay(self):
    self._graph.replay()
    return self._graph_out


class FusedFilterChain:
  """The same chain as FilterChain, forward AND backward in ONE kernel launch (exp_filter_chain_fwd_bwd):
  x and dL/dx_N are read once, x_N and dL/dx_0 written once, the N-1 intermediate images never leave the
  SM (parked in shared memory) -- 48 B/pixel of HBM traffic for the whole chain instead of 60 N.  Usable
  when all N filters and their regressor inputs are known before the first step (this benchmark chain, a
  recorded episode); the agent's rollout needs FilterChain / the per-step ops.

  Regressor inputs and gradients use the kernel's native layout: one [N, B, 24] tensor each."""

  def __init__(self, ids, batch, device, static_chain=True):
    """ids: list of N <= 8 entries, each an int (uniform) or a CUDA int32 tensor [B].  When every entry is an int
    the chain is the same for all images: it runs through exp_filter_chain_fwd_bwd_uniform, which has compile-time
    instantiations for known sequences (the cfg.filters order E,G,W,S+,T,Ct,BW,C); static_chain=False keeps the
    run-time kernel (A/B switch)."""
    self.n = len(ids)
    self.nk = [ops.NUM_PARAMS_ALL[f] if isinstance(f, int) else ops.PSTRIDE for f in ids]
    self.uniform = [int(f) for f in ids] if all(isinstance(f, int) for f in ids) else None
    self.static_chain = static_chain
    self.ids = torch.empty(self.n, batch, dtype=torch.int32, device=device)
    for k, f in enumerate(ids):
      self.ids[k] = f if isinstance(f, int) else f.to(device=device, dtype=torch.int32)
    self.logits = torch.zeros(self.n, batch, ops.PSTRIDE, device=device)
    self.glogits = torch.zeros(self.n, batch, ops.PSTRIDE, device=device)
    self._graph = None

  def set_logits(self, logits_list):
    """logits_list[k]: [B, >= n_k] raw regressor inputs of step k -> packed into self.logits."""
    for k, l in enumerate(logits_list):
      n = min(self.nk[k], l.shape[1])
      self.logits[k, :, :n].copy_(l[:, :n])

  def forward_backward(self, x, gout, need_output=True, need_input_grad=True, y_out=None, gx_out=None):
    """Returns (x_N or None, dL/dx_0 or None, dL/dlogits [N,B,24]) for the logits in self.logits."""
    y, gx, gl = ops.filter_chain_fwd_bwd(x, gout, self.logits, self.uniform if self.uniform is not None else self.ids,
                                         need_y=need_output, need_gx=need_input_grad, logits=True, y_out=y_out,
                                         gx_out=gx_out, gparams_out=self.glogits, static_chain=self.static_chain)
    return y, gx, gl

  def glogits_list(self):
    return [self.glogits[k, :, :n] for k, n in enumerate(self.nk)]

  def capture(self, x, gout):
    """Record forward_backward(x, gout) into a CUDA graph over static buffers (x, gout, self.logits)."""
    y, gx = torch.empty_like(x), torch.empty_like(x)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
      for _ in range(2):
        self.forward_backward(x, gout, y_out=y, gx_out=gx)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    self._graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(self._graph):
      self._graph_out = self.forward_backward(x, gout, y_out=y, gx_out=gx)
    return self._graph_out

  def replay(self):
    self._graph.replay()
    return self._graph_out


class HostPipelinedChain:
  """chain fwd+bwd for batches that live in HOST (pinned) memory: the batch is cut into `chunks`
  sub-batches and H2D copy / compute / D2H copy of consecutive sub-batches overlap on three CUDA
  streams (PCIe is full duplex), so a step costs ~max(H2D, compute, D2H) instead of their sum.

  This is the public end-to-end entry point bench.py's `e2e` leg times."""

  def __init__(self, ids, batch, height, width, device, chunks=4, variant=ops.VARIANT_AUTO, fused=False):
    """fused=True: every sub-batch runs the whole chain forward+backward as ONE kernel (FusedFilterChain)
    instead of 2N per-step launches; results agree with the per-step path to reduction order."""
    assert batch % chunks == 0
    self.ids, self.chunks, self.cb = list(ids), chunks, batch // chunks
    self.device = device
    self.fused = fused
    self.nk = [ops.NUM_PARAMS_ALL[f] if isinstance(f, int) else ops.PSTRIDE for f in self.ids]
    cb, n = self.cb, len(self.ids)
    if fused:
      self.sub = [FusedFilterChain(ids, cb, device) for _ in range(chunks)]
      # regressor inputs of all sub-batches in the kernel's layout, packed once per step: [chunks][N][cb][24]
      self._packed = torch.zeros(chunks, n, cb, ops.PSTRIDE, device=device)
      for c, ch in enumerate(self.sub):
        ch.logits = self._packed[c]
        ch.x, ch.y, ch.gx = (torch.empty(cb, height, width, 3, device=device) for _ in range(3))
      self.h_glog = torch.empty(chunks, n * cb * ops.PSTRIDE).pin_memory()
    else:
      self.sub = [FilterChain(ids, variant=variant) for _ in range(chunks)]
      # the parameter gradients of all steps of a sub-batch live in ONE flat device buffer so that they
      # leave in one D2H copy per sub-batch (a copy per step costs more in launch overhead than in bytes)
      self.off = [0]
      for k in self.nk:
        self.off.append(self.off[-1] + cb * k)
      for ch in self.sub:
        ch.input_buffer((cb, height, width, 3), device)
        ch._glog_flat = torch.zeros(self.off[-1], device=device)
        ch._glog = [ch._glog_flat[self.off[k]:self.off[k + 1]].view(cb, k_n) for k, k_n in enumerate(self.nk)]
      self.h_glog = torch.empty(chunks, self.off[-1]).pin_memory()
    self._scatter_to = None
    self._gy = None                       # per-chunk device staging of dL/dx_N when it arrives from the host
    self.s_in, self.s_cmp, self.s_out = (torch.cuda.Stream(device=device) for _ in range(3))
    self.ev_in = [torch.cuda.Event() for _ in range(chunks)]
    self.ev_cmp = [torch.cuda.Event() for _ in range(chunks)]
    self.ev_out = [torch.cuda.Event() for _ in range(chunks)]
    self._first = True

  def step(self, hx, logits_list, gout, hy, hglogits, wait=True, hgx=None):
    """hx, hy: pinned host [B,H,W,3]; logits_list[k]: device [B,n_k]; gout = dL/dx_N: device [B,H,W,3] or
    pinned host (then it is copied in per sub-batch like hx); hglogits[k]: pinned host [B,n_k]; hgx
    (optional): pinned host [B,H,W,3] that receives dL/dx_0.  Blocks until the results are in host memory, unless
    wait=False: then the step is only enqueued (call wait() before reading hy / hglogits) and the
    H2D copies of the next step overlap this step's compute and D2H -- consecutive steps keep both
    PCIe directions busy without a fill / drain bubble per step."""
    cb = self.cb
    if [int(l.shape[1]) for l in logits_list] != self.nk:
      raise ValueError("logits_list[k] must be [B, %s] (one column per filter parameter)" % self.nk)
    host_g = not gout.is_cuda
    if host_g and self._gy is None:
      self._gy = [torch.empty(cb, *hx.shape[1:], device=self.device) for _ in range(self.chunks)]
    cur = torch.cuda.current_stream()
    for s in (self.s_in, self.s_cmp, self.s_out):
      s.wait_stream(cur)
    if self.fused:
      with torch.cuda.stream(self.s_cmp):               # after the previous step's kernels (same stream)
        for k, (l, n) in enumerate(zip(logits_list, self.nk)):
          self._packed[:, k, :, :n].copy_(l.view(self.chunks, cb, n))
    for c, ch in enumerate(self.sub):
      sl = slice(c * cb, (c + 1) * cb)
      with torch.cuda.stream(self.s_in):
        if not self._first:
          self.s_in.wait_event(self.ev_cmp[c])          # previous step's compute on this buffer is done
        (ch.x if self.fused else ch._acts[0]).copy_(hx[sl], non_blocking=True)
        if host_g:
          self._gy[c].copy_(gout[sl], non_blocking=True)
        self.ev_in[c].record(self.s_in)
      with torch.cuda.stream(self.s_cmp):
        self.s_cmp.wait_event(self.ev_in[c])
        if not self._first:
          self.s_cmp.wait_event(self.ev_out[c])         # previous step's D2H of this chunk's outputs is done
        gy_c = self._gy[c] if host_g else gout[sl]
        if self.fused:
          y, gx, glog_flat = ch.forward_backward(ch.x, gy_c, y_out=ch.y, gx_out=ch.gx)
          glog_flat = glog_flat.view(-1)
        else:
          y = ch.forward_resident([l[sl] for l in logits_list])
          gx, _ = ch.backward(gy_c, need_input_grad=True)
          glog_flat = ch._glog_flat
        self.ev_cmp[c].record(self.s_cmp)
      with torch.cuda.stream(self.s_out):
        self.s_out.wait_event(self.ev_cmp[c])
        hy[sl].copy_(y, non_blocking=True)
        if hgx is not None:
          hgx[sl].copy_(gx, non_blocking=True)
        self.h_glog[c].copy_(glog_flat, non_blocking=True)
        self.ev_out[c].record(self.s_out)
    self._first = False
    self._scatter_to = hglogits
    if wait:
      self.wait()

  def wait(self):
    """Block until every enqueued step has delivered its results to host memory (hy of every step;
    hglogits of the most recent step)."""
    cur = torch.cuda.current_stream()
    cur.wait_stream(self.s_out)
    cur.synchronize()
    if self._scatter_to is not None:
      cb = self.cb
      if self.fused:
        hg = self.h_glog.view(self.chunks, len(self.ids), cb, ops.PSTRIDE)
        for k, (h, n) in enumerate(zip(self._scatter_to, self.nk)):
          h.view(self.chunks, cb, n).copy_(hg[:, k, :, :n])
      else:
        for k, (h, n) in enumerate(zip(self._scatter_to, self.nk)):
          h.view(self.chunks, cb, n).copy_(self.h_glog[:, self.off[k]:self.off[k + 1]].view(self.chunks, cb, n))
      self._scatter_to = None


def shard_plan(global_batch, height, rank, world):
  """How a global batch of `global_batch` images of `height` rows is split over `world` ranks.

  * global_batch >= world (and divisible): by IMAGE -- rank r owns global_batch / world whole images; filters and
    per-image parameter gradients need no communication.
  * global_batch < world (world divisible by it): by ROWS -- world / global_batch ranks share one image, each owns a
    block of height / (world / global_batch) rows.  Every filter is per pixel, so the only cross-rank term is the
    per-image parameter gradient: each rank produces the partial sums of its rows and ONE small all-reduce
    (<= 8 steps x 24 floats per image) adds them.
  Returns dict(mode, images=[first, count], rows=[first, count], group=ranks sharing my image)."""
  if global_batch >= world:
    if global_batch % world:
      raise ValueError("global batch %d is not a multiple of %d ranks" % (global_batch, world))
    per = global_batch // world
    return dict(mode="image", images=(rank * per, per), rows=(0, height), group=[rank])
  if world % global_batch:
    raise ValueError("%d ranks cannot share %d images evenly" % (world, global_batch))
  k = world // global_batch                     # ranks per image
  if height % k:
    raise ValueError("image height %d is not a multiple of %d row blocks" % (height, k))
  img, blk = rank // k, rank % k
  rows = height // k
  return dict(mode="rows", images=(img, 1), rows=(blk * rows, rows), group=list(range(img * k, (img + 1) * k)))


class ShardedFilterChain:
  """The whole-chain forward + backward of a GLOBAL batch spread over the ranks of a process group (BASELINE
  configs[4]: 4K frames, batch 1 -> 32, on 1 / 2 / 4 / 8 GPUs).  Each rank holds only its shard (whole images, or a
  block of rows of one image when there are fewer images than ranks) and runs FusedFilterChain on it; with row
  sharding the parameter gradients of an image are partial sums over each rank's rows -- they are linear in the
  per-pixel terms (finalize + regressor chain rule included), so one all-reduce of the [S, global_batch, 24] gradient
  tensor completes them (every rank writes only its own image's rows of that tensor)."""

  def __init__(self, ids, global_batch, height, width, device, rank=0, world=1):
    self.plan = shard_plan(global_batch, height, rank, world)
    self.global_batch, self.world, self.rank = global_batch, world, rank
    self.local_batch = self.plan["images"][1]
    self.local_rows = self.plan["rows"][1]
    self.width = width
    self.chain = FusedFilterChain(ids, self.local_batch, device)
    self.nk = self.chain.nk
    self._gl_global = torch.zeros(len(ids), global_batch, ops.PSTRIDE, device=device) if self.plan["mode"] == "rows" else None

  def set_logits(self, logits_global):
    """logits_global[k]: [global_batch, n_k] (replicated on every rank: the parameters are per image, tiny)."""
    b0, nb = self.plan["images"]
    self.chain.set_logits([l[b0:b0 + nb] for l in logits_global])

  def forward_backward(self, x_shard, gy_shard, y_out=None, gx_out=None):
    """x_shard, gy_shard: [local_batch, local_rows, W, 3] -- this rank's part.  Returns (y_shard, gx_shard, glogits):
    glogits [S, local_batch, 24] for image sharding, [S, global_batch, 24] (complete, identical on the ranks) for
    row sharding."""
    import torch.distributed as dist
    y, gx, gl = self.chain.forward_backward(x_shard, gy_shard, y_out=y_out, gx_out=gx_out)
    if self.plan["mode"] == "image":
      return y, gx, gl
    self._gl_global.zero_()
    self._gl_global[:, self.plan["images"][0]] = gl[:, 0]
    dist.all_reduce(self._gl_global)                       # <= 8 x 24 floats per image: the only exchange of the path
    return y, gx, self._gl_global
