#!/usr/bin/env python
"""bench.py -- throughput of the Exposure hot path on B200 (driver contract: one JSON line).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                  [--workload chain8] [--batch B] [--size S]

Workload `chain8` (default; BASELINE.json configs[1]): the 8-filter chain
E,G,W,S+,T,Ct,BW,C applied in sequence to a batch of B x S x S x 3 fp32 linear-RGB images,
forward + backward (image gradient and per-image parameter gradients), B=64, S=512 per GPU.
A "step" is one such pass over one batch.  Multi-GPU: the batch shards by image, one
process per GPU, no data-path collective (weak scaling: 64 images per GPU).

Reported:
  value     images/s, whole job, inputs resident in HBM (CUDA events, max over ranks)
  e2e       images/s through the public API with HOST (pinned) buffers: H2D of the batch and
            D2H of the filtered batch + parameter gradients inside the timed region
  roofline  the dominant kernel's achieved algorithmic GB/s vs MEASURED_PEAKS.json
  cpu_baseline  the CPU restatement of the reference TF graph (oracle/, torch fp32,
            op by op + autograd) on a bounded sample, all host threads
`--impl reference` times only that CPU restatement (the TF-1.6 reference cannot run here).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CHAIN_IDS = [0, 1, 2, 3, 4, 5, 6, 7]         # E,G,W,S+,T,Ct,BW,C = cfg.filters order
FWD_B, BWD_B = 24, 36                        # algorithmic bytes / pixel / step


def parse_args():
  ap = argparse.ArgumentParser()
  ap.add_argument("--gpus", type=int, default=1)
  ap.add_argument("--steps", type=int, default=400)
  ap.add_argument("--warmup", type=int, default=10)
  ap.add_argument("--impl", default="native", choices=["native", "reference"])
  ap.add_argument("--workload", default="chain8", choices=["chain8", "train", "eval"])
  ap.add_argument("--batch", type=int, default=64, help="images per GPU")
  ap.add_argument("--size", type=int, default=512)
  ap.add_argument("--height", type=int, default=0, help="chain8: image height (default --size); 2160 for the 4K config")
  ap.add_argument("--width", type=int, default=0, help="chain8: image width (default --size); 3840 for the 4K config")
  ap.add_argument("--variant", type=int, default=0, help="0 auto, 1 direct, 2 tma (exposure_b200.h)")
  ap.add_argument("--chain-impl", default="fused", choices=["fused", "steps"],
                  help="chain8: 'fused' = the whole chain fwd+bwd as ONE kernel per step (exp_filter_chain_fwd_bwd), "
                       "'steps' = one fused kernel per filter step and direction (2N launches)")
  ap.add_argument("--e2e-chunks", type=int, default=0, help="chain8 e2e leg: sub-batches in flight (0 = 8 if it divides the batch)")
  ap.add_argument("--no-cpu-baseline", action="store_true")
  ap.add_argument("--no-graphs", action="store_true", help="train workload: launch kernels eagerly instead of CUDA graphs")
  return ap.parse_args()


def host_threads():
  """Host threads this process may really use: min(affinity mask, cgroup cpu quota)."""
  n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
  try:
    with open("/sys/fs/cgroup/cpu.max") as fh:
      quota, period = fh.read().split()
    if quota != "max":
      n = max(1, min(n, int(float(quota) / float(period))))
  except Exception:
    pass
  return n


def peaks():
  p = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(p):
    with open(p) as fh:
      d = json.load(fh)
    return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
  return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------
# CPU baseline: the oracle = CPU restatement of the reference TF graph (NOT TensorFlow)
# ------------------------------------------------------------------------------------------
def cpu_chain_step(x, logits, gout):
  """One chain8 fwd+bwd the way the reference executes: op by op, unfused, tf.gradients ==
  autograd of the forward graph (oracle/filters.py)."""
  import torch
  from oracle import filters as F
  xs = x.clone().requires_grad_(True)
  lg = [l.clone().requires_grad_(True) for l in logits]
  y = F.chain_fwd(CHAIN_IDS, xs, lg)[-1]
  torch.autograd.grad(y, [xs] + lg, grad_outputs=gout)
  return y


def cpu_baseline(size, budget_s=20.0, batch=2):
  import torch
  from oracle import filters as F
  cores = host_threads()
  torch.set_num_threads(cores)
  x = F.synth_images(batch, size, size, seed=1234, stress=False)
  logits = [F.synth_logits(f, batch) for f in CHAIN_IDS]
  gout = torch.ones_like(x)
  cpu_chain_step(x, logits, gout)                      # warm-up
  times = []
  t_end = time.time() + budget_s
  while len(times) < 10 and (time.time() < t_end or len(times) < 2):
    t0 = time.time()
    cpu_chain_step(x, logits, gout)
    times.append(time.time() - t0)
  med = statistics.median(times)
  return {
      "value": batch / med, "unit": "images/s", "cores": cores, "kind": "port",
      "sample": "chain8 fwd+bwd on %dx%dx%dx3 fp32, median of %d reps (min %.3fs); CPU restatement of the "
                "reference TF graph (oracle/filters.py, torch-CPU op-by-op + autograd), not TensorFlow" %
                (batch, size, size, len(times), min(times)),
  }


def run_reference(args):
  """--impl reference: the reference's CPU path (oracle port) on the host cores; rank 0 only."""
  rank = int(os.environ.get("RANK", "0"))
  if rank != 0:
    return
  import torch
  from oracle import filters as F
  if args.workload == "train":
    iteration, sb, cores = _cpu_train_iteration(args.batch)
    for _ in range(max(1, min(args.warmup, 2))):
      iteration()
    t0 = time.time()
    for _ in range(args.steps):
      iteration()
    dt = time.time() - t0
    val = sb * args.steps / dt
    sample = ("each step = 1 generator+value step + 5 critic steps on a bounded sample of %d of the %d images; CPU "
              "restatement of the reference TF graph (oracle/train_step.py port, torch-CPU autograd), %d threads" %
              (sb, args.batch, cores))
    print(json.dumps({
        "impl": "reference", "metric": "images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": train_config(args),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return
  cores = host_threads()
  torch.set_num_threads(cores)
  sample_b = 2
  x = F.synth_images(sample_b, args.size, args.size, seed=1234, stress=False)
  logits = [F.synth_logits(f, sample_b) for f in CHAIN_IDS]
  gout = torch.ones_like(x)
  for _ in range(max(1, min(args.warmup, 2))):
    cpu_chain_step(x, logits, gout)
  t0 = time.time()
  for _ in range(args.steps):
    cpu_chain_step(x, logits, gout)
  dt = time.time() - t0
  val = sample_b * args.steps / dt
  sample = ("each step = chain8 fwd+bwd on a bounded sample of %d of the %d images (%dx%dx3 fp32); CPU "
            "restatement of the reference TF graph (oracle port), torch-CPU, %d threads" %
            (sample_b, args.batch, args.size, args.size, cores))
  print(json.dumps({
      "impl": "reference", "metric": "images/sec", "value": val, "unit": "images/s", "n_gpus": args.gpus,
      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
      "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": workload_config(args),
      "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
      "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
      "gpu_launches": 0,
  }))


def workload_config(args):
  return {
      "workload": "chain8: 8-filter chain E,G,W,S+,T,Ct,BW,C fwd+bwd (BASELINE.json configs[1])",
      "batch_per_gpu": args.batch, "height": args.height or args.size, "width": args.width or args.size, "channels": 3,
      "filters": "E,G,W,S+,T,Ct,BW,C", "parallelism": "dp%d (batch sharded by image, no data-path collective)" % args.gpus,
      "l2_policy": "working set (%s x %.0f MB) exceeds the 126 MB L2; no flush needed"
                   % ("x, dL/dy, y, dL/dx: 4" if args.chain_impl == "fused" else "9 activations + 2 gradient buffers: 11",
                      args.batch * (args.height or args.size) * (args.width or args.size) * 12 / 1e6),
  }


# ------------------------------------------------------------------------------------------
class ClockSampler:
  """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
  Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
       "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
       "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

  def __init__(self, gpu_index):
    self.idx = gpu_index
    self.proc = None
    self.path = None

  def start(self):
    try:
      fd, self.path = tempfile.mkstemp(suffix=".csv")
      os.close(fd)
      self.fh = open(self.path, "w")
      self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits", "-lms", "25"], stdout=self.fh,
                                   stderr=subprocess.DEVNULL)
    except Exception:
      self.proc = None

  def stop(self, t0=None, t1=None):
    """Summary of the samples taken in the wall-clock window [t0, t1] (the timed region);
    falls back to every sample taken while the bench was under load when the window holds
    fewer than 3."""
    if self.proc is None:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    self.proc.terminate()
    try:
      self.proc.wait(timeout=5)
    except Exception:
      self.proc.kill()
    self.fh.close()
    import datetime
    rows = []
    for line in open(self.path):
      f = [t.strip() for t in line.split(",")]
      if len(f) < 9:
        continue
      try:
        ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
        rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[5:9]))
      except ValueError:
        continue
    os.unlink(self.path)
    window = "timed region"
    sel = [r for r in rows if t0 is not None and t0 <= r[0] <= t1]
    if len(sel) < 3:
      sel, window = rows, "whole loaded run (timed region too short for 3 samples)"
    if not sel:
      return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
    reasons = set()
    for r in sel:
      for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
        if v.lower().startswith("active"):
          reasons.add(name)
    return {"sm_mhz": statistics.median(r[1] for r in sel), "sm_max_mhz": max(r[2] for r in sel),
            "reasons": sorted(reasons), "samples": len(sel), "power_w_max": max(r[3] for r in sel),
            "window": window}


def run_native(args):
  import torch
  import torch.distributed as dist
  from exposure_b200 import ops
  from exposure_b200.chain import FilterChain

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  B, S = args.batch, args.size
  H, W = args.height or S, args.width or S

  # synthetic linear-RGB batch, generated on the device for the resident leg
  g = torch.Generator(device=dev).manual_seed(1234 + rank)
  chain = FilterChain(CHAIN_IDS, variant=args.variant)
  x0 = chain.input_buffer((B, H, W, 3), dev)
  x0.copy_(torch.exp(torch.randn(B, H, W, 3, device=dev, generator=g) - 3.2).clamp_(0, 4))
  stress = torch.rand(B, H, W, 3, device=dev, generator=g)
  x0.copy_(torch.where(stress < 0.01, 1 + 3 * torch.rand(B, H, W, 3, device=dev, generator=g), x0))
  del stress
  gl = torch.Generator().manual_seed(4321)
  logits = [torch.randn(B, ops.NUM_PARAMS[f], generator=gl).to(dev) for f in CHAIN_IDS]
  gout = torch.randn(B, H, W, 3, device=dev, generator=g)

  def step():
    chain.forward_resident(logits)
    return chain.backward(gout)

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  for _ in range(max(args.warmup, 3)):
    step()
  barrier()

  # ---- timed region (device resident) -----------------------------------------------------
  ops.event_log = []
  l0 = ops.launch_count
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  barrier()
  wall0 = time.time()
  e0.record()
  for _ in range(args.steps):
    step()
  e1.record()
  barrier()
  wall1 = time.time()
  elapsed_ms = e0.elapsed_time(e1)
  launches = ops.launch_count - l0
  log, ops.event_log = ops.event_log, None
  eager_ms = elapsed_ms
  graph_info = None
  if not args.no_graphs:
    # the same K steps replayed from ONE CUDA graph (2N kernel nodes): this is the timed region
    # `value` reports; the eager loop above is the instrumented pass for the per-kernel table
    chain.capture(logits, gout)
    for _ in range(3):
      chain.replay()
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
      chain.replay()
    e1.record()
    barrier()
    wall1 = time.time()
    elapsed_ms = e0.elapsed_time(e1)
    launches = args.steps * chain.graph_launches
    graph_info = {"nodes_per_step": chain.graph_launches, "eager_ms_per_step": eager_ms / args.steps}
  t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  elapsed_ms = float(t.item())
  value = world * B * args.steps / (elapsed_ms / 1e3)

  # ---- per-kernel roofline from the events recorded inside the (eager) timed loop -----------
  per = {}
  for name, nbytes, a, b in log:
    d = per.setdefault(name, {"ms": 0.0, "bytes": 0, "n": 0})
    d["ms"] += a.elapsed_time(b); d["bytes"] += nbytes; d["n"] += 1
  peak, peak_src = peaks()
  kernels = []
  for name, d in sorted(per.items(), key=lambda kv: -kv[1]["ms"]):
    kernels.append({"kernel": name, "launches": d["n"], "avg_ms": d["ms"] / d["n"],
                    "algorithmic_bytes_per_launch": d["bytes"] // d["n"],
                    "achieved_gbs": d["bytes"] / d["ms"] / 1e6, "frac": d["bytes"] / d["ms"] / 1e6 / peak})
  tot_ms = sum(d["ms"] for d in per.values())
  tot_bytes = sum(d["bytes"] for d in per.values())
  dom = kernels[0]
  # DRAM traffic per launch of the dominant kernel from the committed `ncu --set full` capture of this
  # same command (profiles/filter_traffic.json, written by profiles/summarize.py traffic); only valid
  # for the shape it was captured on
  traffic, traffic_src = None, None
  try:
    tj = json.load(open(os.path.join(ROOT, "profiles", "filter_traffic.json")))
    if tj.get("config") == "chain8 %dx%dx%dx3 fp32" % (B, H, W) and dom["kernel"] in tj["kernels"]:
      traffic = tj["kernels"][dom["kernel"]]["traffic_bytes"]
      traffic_src = "profiles/filter_traffic.json (%s): dram__bytes_read.sum + dram__bytes_write.sum of one launch" % tj["source"]
  except (OSError, ValueError, KeyError):
    pass
  roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved_gbs"], "peak": peak,
              "unit": "GB/s", "frac": dom["frac"], "traffic": traffic, "traffic_source": traffic_src,
              "peak_source": peak_src,
              "share_of_step": dom["avg_ms"] * dom["launches"] / eager_ms if eager_ms else None,
              "chain": {"achieved": tot_bytes / tot_ms / 1e6, "frac": tot_bytes / tot_ms / 1e6 / peak,
                        "algorithmic_bytes_per_step": B * H * W * (FWD_B + BWD_B) * len(CHAIN_IDS),
                        "kernel_ms_per_step": tot_ms / args.steps,
                        "step_frac": B * H * W * (FWD_B + BWD_B) * len(CHAIN_IDS) / (elapsed_ms / args.steps) / 1e6 / peak},
              "note": "kernel durations: CUDA events around every launch of the eager K-step loop; `value` and "
                      "chain.step_frac: the same K steps replayed from one CUDA graph" if graph_info else
                      "kernel durations: CUDA events around every launch inside the timed region",
              "cuda_graph": graph_info,
              "kernels": kernels}

  # ---- the whole chain as ONE kernel per step (exp_filter_chain_fwd_bwd) ----------------------
  fused_on = args.chain_impl == "fused"
  if fused_on:
    from exposure_b200.chain import FusedFilterChain
    per_step = {"value": value, "ms_per_step": elapsed_ms / args.steps, "gpu_launches": launches, "roofline": roofline}
    y_steps = chain._acts[-1]
    fz = FusedFilterChain(CHAIN_IDS, B, dev)
    fz.set_logits(logits)
    fy, fgx = torch.empty_like(x0), torch.empty_like(x0)
    for _ in range(max(args.warmup, 3)):
      fz.forward_backward(x0, gout, y_out=fy, gx_out=fgx)
    barrier()
    # same inputs, same outputs: largest relative deviation from the per-step kernels' result
    agree = float(((fy - y_steps).abs() / y_steps.abs().clamp_min(1e-3)).max())
    ops.event_log = []
    for _ in range(min(args.steps, 50)):
      fz.forward_backward(x0, gout, y_out=fy, gx_out=fgx)
    barrier()
    flog, ops.event_log = ops.event_log, None
    f_launch_ms = sum(a.elapsed_time(b) for _, _, a, b in flog) / len(flog)
    l0 = ops.launch_count
    barrier()
    wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
      fz.forward_backward(x0, gout, y_out=fy, gx_out=fgx)
    e1.record()
    barrier()
    wall1 = time.time()
    elapsed_ms = e0.elapsed_time(e1)
    launches = ops.launch_count - l0
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
      dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    value = world * B * args.steps / (elapsed_ms / 1e3)
    algo = B * H * W * (FWD_B + BWD_B) * len(CHAIN_IDS)       # 60 B/pixel/step, N steps
    actual = B * H * W * 48                                   # what the fused kernel moves: x, gy in; y, gx out
    ftraffic, fsrc = None, None
    try:
      tj = json.load(open(os.path.join(ROOT, "profiles", "filter_traffic.json")))
      if tj.get("config") == "chain8 %dx%dx%dx3 fp32" % (B, H, W) and "filter_chain_fwd_bwd" in tj["kernels"]:
        ftraffic = tj["kernels"]["filter_chain_fwd_bwd"]["traffic_bytes"]
        fsrc = "profiles/filter_traffic.json (%s): dram__bytes_read.sum + dram__bytes_write.sum of one launch" % tj["source"]
    except (OSError, ValueError, KeyError):
      pass
    roofline = {"bound": "hbm", "kernel": "filter_chain_fwd_bwd", "achieved": algo / f_launch_ms / 1e6, "peak": peak,
                "unit": "GB/s", "frac": algo / f_launch_ms / 1e6 / peak, "traffic": ftraffic, "traffic_source": fsrc,
                "peak_source": peak_src, "share_of_step": f_launch_ms * args.steps / elapsed_ms,
                "algorithmic_bytes_per_launch": algo, "launch_ms": f_launch_ms,
                "hbm_bytes_moved_per_launch": actual, "hbm_frac_of_bytes_moved": actual / f_launch_ms / 1e6 / peak,
                "fused_vs_per_step_max_rel_dev": agree,
                "note": "`achieved` uses SURVEY 8d's definition of the path's algorithmic bytes (60 B/pixel/step x N steps), "
                        "as it prescribes for the chain-fused variant; the kernel keeps the N-1 intermediate images in "
                        "shared memory and moves only 48 B/pixel for the whole chain, so frac > 1 against the per-step "
                        "definition and the kernel is instruction-issue bound, not HBM bound (hbm_frac_of_bytes_moved; "
                        "ncu summary under profiles/).  per_step = the same chain as 2N per-step kernels (the HBM-bound "
                        "path the agent's rollout uses), measured in the same run.",
                "per_step": per_step}

  # ---- end-to-end leg: host (pinned) buffers through the public API ------------------------
  hx = torch.empty(B, H, W, 3, dtype=torch.float32).pin_memory()
  hx.copy_(x0.cpu())
  hy = torch.empty(B, H, W, 3, dtype=torch.float32).pin_memory()
  hg = [torch.empty(B, ops.NUM_PARAMS[f]).pin_memory() for f in CHAIN_IDS]
  # public host-buffer API: H2D of the batch, chain fwd+bwd, D2H of the filtered batch (net.py:330
  # fetches fake_output every step) and of the parameter gradients, software-pipelined over
  # sub-batches on three streams (exposure_b200/chain.py HostPipelinedChain)
  from exposure_b200.chain import HostPipelinedChain
  n_chunks = args.e2e_chunks if args.e2e_chunks > 0 and B % args.e2e_chunks == 0 else \
      (8 if B % 8 == 0 else (4 if B % 4 == 0 else 1))
  del chain                                          # free the resident chain's activations first
  if fused_on:
    del fz, fy, fgx, y_steps
  torch.cuda.empty_cache()
  pipe = HostPipelinedChain(CHAIN_IDS, B, H, W, dev, chunks=n_chunks, variant=args.variant, fused=fused_on)

  def e2e_step():
    # enqueue only: step i+1's H2D overlaps step i's compute and D2H (every step still copies its
    # whole input batch in and its whole result out inside the timed region)
    pipe.step(hx, logits, gout, hy, hg, wait=False)

  e2e_steps = max(3, min(args.steps, 10))
  for _ in range(2):
    e2e_step()
  pipe.wait()
  barrier()
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(e2e_steps):
    e2e_step()
  torch.cuda.current_stream().wait_stream(pipe.s_out)
  b.record()
  pipe.wait()
  barrier()
  e2e_ms = a.elapsed_time(b)                     # device clock
  t = torch.tensor([e2e_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
  e2e_ms = float(t.item())
  e2e = {"value": world * B * e2e_steps / (e2e_ms / 1e3), "unit": "images/s",
         "h2d_bytes_per_step": hx.numel() * 4,
         "d2h_bytes_per_step": hy.numel() * 4 + sum(h.numel() * 4 for h in hg), "steps": e2e_steps,
         "ms_per_step": e2e_ms / e2e_steps}

  clocks = sampler.stop(wall0, wall1) if rank == 0 else None
  out = None
  if rank == 0:
    out = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(args), chain_impl=args.chain_impl), "clocks": clocks, "e2e": e2e,
        "gpu_launches": launches,
        "roofline": roofline,
    }
    if world == 1 and not args.no_cpu_baseline:
      out["cpu_baseline"] = cpu_baseline(S)
    elif world == 1:
      out["cpu_baseline"] = None
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  if out is not None:
    print(json.dumps(out))


def train_config(args):
  return {
      "workload": "train: 1 generator+value step + 5 WGAN-GP critic steps per iteration, replay memory on device "
                  "(BASELINE.json configs[3]; net.py:307-370)",
      "batch_per_gpu": args.batch, "height": 64, "width": 64, "channels": 3, "filters": "E,G,W,S+,T,Ct,BW,C",
      "giters": 1, "citers": 5,
      "parallelism": "dp%d (batch sharded by image; one NCCL all-reduce per optimizer step)" % args.gpus,
      "l2_policy": "working set is L2 resident by nature (3 MB batches); every iteration draws fresh replay "
                   "batches, dropout masks and noise",
  }


def run_train(args):
  """configs[3]: full train iteration at batch 64x64x64 per GPU, data parallel."""
  import torch
  import torch.distributed as dist
  from exposure_b200 import ops
  from exposure_b200.replay import ReplayMemory, SyntheticProvider
  from exposure_b200.trainer import Trainer, default_cfg

  world = int(os.environ.get("WORLD_SIZE", "1"))
  rank = int(os.environ.get("RANK", "0"))
  local = int(os.environ.get("LOCAL_RANK", "0"))
  if not torch.cuda.is_available():
    raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  cfg = default_cfg()
  cfg.batch_size = args.batch
  cfg.replay_memory_size = 2 * args.batch
  B = args.batch
  t = Trainer(cfg, dev, seed=0)                       # identical initial weights on every rank
  mem = ReplayMemory(cfg, SyntheticProvider(dev, "raw", 100 + rank), SyntheticProvider(dev, "real", 200 + rank), dev,
                     seed=rank)
  t.attach_memory(mem, torch.Generator(device=dev).manual_seed(300 + rank))

  def barrier():
    if world > 1:
      dist.barrier()
    torch.cuda.synchronize()

  # bootstrap like iteration 0 of net.py:318-328 (generator steps with lr 0 until terminated
  # records exist), shortened: 2 * test_steps generator steps are enough to terminate records
  def note(msg):
    if rank == 0:
      print("[bench train] " + msg, file=sys.stderr, flush=True)

  t.train_iteration(0, giters=2 * cfg.test_steps + 2, citers=1)
  torch.cuda.synchronize()
  note("bootstrap done")
  if not args.no_graphs:
    t.enable_graphs(B)                              # each step = one CUDA-graph replay
    torch.cuda.synchronize()
    note("graphs captured (optimizer inside the graph: %s)" % t._graph_apply)
  sampler = ClockSampler(local)
  if rank == 0:
    sampler.start()
  it = 1
  for _ in range(max(args.warmup, 3)):
    t.train_iteration(it, giters=1, citers=5)
    it += 1
  barrier()
  note("warm-up done")
  ops.event_log = []
  l0 = ops.launch_count
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  wall0 = time.time()
  e0.record()
  last = None
  for _ in range(args.steps):
    last = t.train_iteration(it, giters=1, citers=5)
    it += 1
  e1.record()
  barrier()
  wall1 = time.time()
  elapsed_ms = e0.elapsed_time(e1)
  launches = ops.launch_count - l0
  log, ops.event_log = ops.event_log, None
  graphs = getattr(t, "_ggraph", None) is not None
  roof_note = "per-kernel CUDA events recorded inside the timed region"
  if graphs:
    # kernels replayed from a CUDA graph are not launched through Python: count them from the
    # capture, and time the kernel families in a short eager (non-graph) pass right after
    launches = args.steps * (t.graph_launches["generator"] + 5 * t.graph_launches["critic"])
    saved = (t._ggraph, t._cgraph)
    t._ggraph = t._cgraph = None
    ops.event_log = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    n_probe = 3
    for _ in range(n_probe):
      t.train_iteration(it, giters=1, citers=5)
      it += 1
    ev1.record()
    torch.cuda.synchronize()
    log, ops.event_log = ops.event_log, None
    t._ggraph, t._cgraph = saved
    probe_ms = ev0.elapsed_time(ev1)
    roof_note = ("timed region replays CUDA graphs; kernel families timed with CUDA events in %d eager iterations "
                 "right after it (%.2f ms/iteration eager)" % (n_probe, probe_ms / n_probe))
  tt = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
  elapsed_ms = float(tt.item())
  value = world * B * args.steps / (elapsed_ms / 1e3)

  per = {}
  for name, flops, a, b in log:
    d = per.setdefault(name, {"ms": 0.0, "flops": 0, "n": 0})
    d["ms"] += a.elapsed_time(b); d["flops"] += flops; d["n"] += 1
  kernels = [{"kernel": k, "launches": d["n"], "avg_ms": d["ms"] / d["n"], "tflops": d["flops"] / d["ms"] / 1e9}
             for k, d in sorted(per.items(), key=lambda kv: -kv[1]["ms"])]
  peak_tf = 1389.9
  pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
  if os.path.exists(pk):
    peak_tf = float(json.load(open(pk)).get("bf16_tflops_sustained", peak_tf))
  tot_ms = sum(d["ms"] for d in per.values()) or 1.0
  tot_fl = sum(d["flops"] for d in per.values())
  backend = os.environ.get("EXPOSURE_GEMM_BACKEND", "0")
  kname = {"0": "tma_gemm_kernel family (conv fprop/dgrad/wgrad: TMA-fed tcgen05 kind::tf32, 3xTF32 split, TMEM "
                "accumulators, cluster split-K) + CUDA-core FC / small-N kernels",
           "4": "tma_gemm_kernel family (TMA-fed tcgen05, 3xTF32) + CUDA-core FC / small-N kernels",
           "2": "tc_gemm_kernel family (register-gather tcgen05, 3xTF32)",
           "3": "tc_gemm_ws_kernel family (warp-specialised register-gather tcgen05, 3xTF32)",
           "1": "gemm_kernel family (exact-fp32 CUDA-core engine)"}.get(backend, backend)
  n_timed_iters = 3 if graphs else args.steps
  roofline = {"bound": "tensor", "kernel": kname, "gemm_backend": {"0": "auto (tcgen05-tma where supported)", "1": "cuda-cores", "2": "tcgen05", "3": "tcgen05-ws",
                                                                      "4": "tcgen05-tma"}.get(backend, backend),
              "achieved": tot_fl / tot_ms / 1e9, "peak": peak_tf, "unit": "TFLOP/s", "frac": tot_fl / tot_ms / 1e9 / peak_tf,
              "traffic": None, "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS bf16; the fp32-accurate "
                                              "3xTF32 path can reach at most 1/6 of it)",
              "share_of_step": (tot_ms / n_timed_iters) / (elapsed_ms / args.steps), "note": roof_note,
              "kernels": kernels}

  # end-to-end leg: fresh RAW and real batches come from pinned host memory every step, the
  # filtered batch and the losses go back to the host (what net.py:330-342 does per sess.run)
  hraw = torch.empty(B, 64, 64, 3).pin_memory(); hraw.copy_(mem.fake_dataset.get_next_batch(B).cpu())
  hreal = torch.empty(B, 64, 64, 3).pin_memory(); hreal.copy_(mem.real_dataset.get_next_batch(B).cpu())
  hout = torch.empty(B, 64, 64, 3).pin_memory()
  hloss = torch.empty(4).pin_memory()

  class HostProvider:
    def __init__(self, h):
      self.h = h
    def get_next_batch(self, n):
      return self.h[:n].to(dev, non_blocking=True)

  mem.fake_dataset, mem.real_dataset = HostProvider(hraw), HostProvider(hreal)
  h2d = [0]

  def e2e_iter():
    nonlocal it
    out = t.train_iteration(it, giters=1, citers=5)
    it += 1
    hloss.copy_(torch.stack([out["g_loss"], out["v_loss"], out["emd"], out["critic_gradient_norm"]]), non_blocking=True)
    hout.copy_(mem.images[:B], non_blocking=True)
    torch.cuda.current_stream().synchronize()

  for _ in range(2):
    e2e_iter()
  barrier()
  n_e2e = max(3, min(args.steps, 20))
  a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  a.record()
  for _ in range(n_e2e):
    e2e_iter()
  b.record()
  barrier()
  tt = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
  if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
  e2e_ms = float(tt.item())
  e2e = {"value": world * B * n_e2e / (e2e_ms / 1e3), "unit": "images/s",
         "h2d_bytes_per_step": 6 * B * 64 * 64 * 3 * 4, "d2h_bytes_per_step": B * 64 * 64 * 3 * 4 + 16,
         "steps": n_e2e, "ms_per_step": e2e_ms / n_e2e,
         "note": "h2d upper bound: 5 real batches + up to 1 fresh RAW batch per iteration"}
  clocks = sampler.stop(wall0, wall1) if rank == 0 else None
  if rank == 0:
    out = {"metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": train_config(args),
           "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
           "losses": {k: (float(v) if v is not None else None) for k, v in last.items()}}
    if world == 1:
      out["cpu_baseline"] = None if args.no_cpu_baseline else cpu_baseline_train(B)
  if world > 1:
    dist.barrier()
    dist.destroy_process_group()
  if rank == 0:
    print(json.dumps(out))


def _cpu_train_iteration(B):
  """The reference's CPU path for one train iteration: oracle generator step + 5 critic steps
  (autograd, all 8 filters + one-hot select), on a bounded sample of the batch.  Returns
  (callable running one iteration, images in the sample, host threads used)."""
  import torch
  from oracle import filters as OF
  from oracle import train_step as OT
  from exposure_b200.trainer import default_cfg
  cores = host_threads()
  torch.set_num_threads(cores)
  cfg = default_cfg()
  sb = min(B, 8)
  g = torch.Generator().manual_seed(0)
  shapes_c = {"critic": 6, "rl_value/critic": 17}
  def make(scope, cin):
    P = {}
    c = cin
    for i, co in enumerate((32, 64, 128, 256)):
      n = "%s/Conv%s" % (scope, "" if i == 0 else "_%d" % i)
      P[n + "/weights"] = torch.randn(4, 4, c, co, generator=g) * 0.05
      P[n + "/biases"] = torch.zeros(co)
      c = co
    return P
  Pc = make("critic", 6); Pv = make("rl_value/critic", 17)
  for P, s in ((Pc, "critic"), (Pv, "rl_value/critic")):
    P[s + "/fully_connected/weights"] = torch.randn(4096, 128, generator=g) * 0.02; P[s + "/fully_connected/biases"] = torch.zeros(128)
    P[s + "/fully_connected_1/weights"] = torch.randn(128, 1, generator=g) * 0.1; P[s + "/fully_connected_1/biases"] = torch.zeros(1)
  Pg = make("generator", 14); Pg.update(make("generator/action_selection", 14))
  for j, n in enumerate(OF.NUM_PARAMS[:8]):
    Pg["generator/filter_%d/fc1/weights" % j] = torch.randn(4096, 128, generator=g) * 0.02
    Pg["generator/filter_%d/fc1/biases" % j] = torch.zeros(128)
    Pg["generator/filter_%d/fc2/weights" % j] = torch.randn(128, n + 6, generator=g) * 0.1
    Pg["generator/filter_%d/fc2/biases" % j] = torch.zeros(n + 6)
  Pg["generator/action_selection/selector_fc1/weights"] = torch.randn(4096, 128, generator=g) * 0.02
  Pg["generator/action_selection/selector_fc1/biases"] = torch.zeros(128)
  Pg["generator/action_selection/selector_fc2/weights"] = torch.randn(128, 8, generator=g) * 0.1
  Pg["generator/action_selection/selector_fc2/biases"] = torch.zeros(8)
  img = OF.synth_images(sb, 64, 64, stress=False)
  real = OF.synth_images(sb, 64, 64, seed=9, stress=False) * 4
  states = torch.zeros(sb, 11)
  noise = torch.rand(sb, generator=g)
  dm = lambda: (torch.rand(sb, 4, 4, 256, generator=g) < 0.5).float() * 2
  alpha = torch.rand(sb, generator=g)

  def iteration():
    out = OT.generator_step(Pg, Pv, Pc, img, states, noise, dm(), dm(), 0.1, cfg)
    for _ in range(5):
      OT.critic_step(Pc, real, out["fake_output"], alpha, cfg)

  return iteration, sb, cores


def cpu_baseline_train(B, budget_s=25.0):
  iteration, sb, cores = _cpu_train_iteration(B)
  iteration()
  times = []
  t_end = time.time() + budget_s
  while len(times) < 5 and (time.time() < t_end or len(times) < 2):
    t0 = time.time()
    iteration()
    times.append(time.time() - t0)
  med = statistics.median(times)
  return {"value": sb / med, "unit": "images/s", "cores": cores, "kind": "port",
          "sample": "1 generator+value step + 5 critic steps on %d of the %d images, median of %d reps; CPU restatement of "
                    "the reference TF graph (oracle/train_step.py, torch-CPU autograd), not TensorFlow" % (sb, B, len(times))}


def run_eval(args):
  """configs[2] (evaluate.py inference): cfg.test_steps policy steps on 64x64 thumbnails + the
  selected filters applied to the (optionally high-resolution) batch by ONE fused kernel."""
  import torch
  from exposure_b200 import ops
  from exposure_b200.evaluate import retouch
  from exposure_b200.trainer import Trainer
  if not torch.cuda.is_available():
    raise SystemExit("bench.py --impl native needs a CUDA device (no CPU fallback)")
  local = int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  B = args.batch
  H, W = args.height or 64, args.width or 64
  t = Trainer(device=dev, seed=0)
  g = torch.Generator(device=dev).manual_seed(77)
  hi = torch.exp(torch.randn(B, H, W, 3, device=dev, generator=g) - 3.2).clamp_(0, 4)
  for _ in range(max(args.warmup, 3)):
    retouch(t, hi, generator=g)
  torch.cuda.synchronize()
  ops.event_log = []
  l0 = ops.launch_count
  e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
  e0.record()
  for _ in range(args.steps):
    out = retouch(t, hi, generator=g)
  e1.record()
  torch.cuda.synchronize()
  ms = e0.elapsed_time(e1)
  log, ops.event_log = ops.event_log, None
  fused = [(a.elapsed_time(b), nb) for name, nb, a, b in log if name.startswith("filter_chain_fwd")]
  peak, peak_src = peaks()
  fms = sum(x[0] for x in fused) / max(1, len(fused))
  fbytes = fused[0][1] if fused else 0
  S = int(out["ids"].shape[0])
  print(json.dumps({
      "metric": "images/sec", "value": B * args.steps / (ms / 1e3), "unit": "images/s", "n_gpus": 1, "steps": args.steps,
      "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
      "vs_baseline": None, "dtype": "f32", "data": "synthetic",
      "config": {"workload": "eval: %d policy steps on 64x64 thumbnails + fused %d-step filter apply on the %dx%d batch "
                             "(BASELINE.json configs[2]; net.py:796-820)" % (S, S, H, W),
                 "batch_per_gpu": B, "height": H, "width": W, "policy_steps": S},
      "gpu_launches": ops.launch_count - l0,
      "roofline": {"bound": "hbm", "kernel": "filter_chain_fwd_kernel (all %d steps in one pass)" % S,
                   "achieved": fbytes / fms / 1e6 if fms else None, "peak": peak, "unit": "GB/s",
                   "frac": fbytes / fms / 1e6 / peak if fms else None, "traffic": None, "peak_source": peak_src,
                   "avg_ms": fms, "algorithmic_bytes_per_launch": fbytes,
                   "note": "24 B/pixel for the whole episode; the unfused schedule moves %d B/pixel" % (24 * S)},
  }))


def main():
  args = parse_args()
  # stdout must carry exactly ONE JSON line: native libraries write banners to fd 1 (NCCL prints
  # "NCCL version ..." there), so fd 1 is pointed at stderr and Python's stdout keeps the real one
  sys.stdout.flush()
  real_stdout = os.dup(1)
  os.dup2(2, 1)
  sys.stdout = os.fdopen(real_stdout, "w")
  if args.impl == "reference":
    run_reference(args)
  elif args.workload == "eval":
    run_eval(args)
  elif args.workload == "train":
    run_train(args)
  else:
    run_native(args)


if __name__ == "__main__":
  main()
