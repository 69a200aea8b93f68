#!/usr/bin/env python
"""GPU timeline of the train iteration from CUPTI activity records (torch.profiler; nsys is not in
the image): per-kernel start/end inside the CUDA-graph replays -> busy time (union of kernel
intervals), idle gaps, concurrency, per-kernel totals, and the same split per graph replay.

  python tools/train_timeline.py [--iters 5] [--batch 64] [--out gpurun_out/timeline.json] [--host-replay]
  torchrun --nproc-per-node N tools/train_timeline.py ...     (data parallel: rank 0 records and writes its own timeline)

A measurement aid only (numbers under a profiler are never bench values)."""
import argparse
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def short(name):
  name = re.sub(r"^void ", "", name)
  name = re.sub(r"\(.*$", "", name)
  name = name.replace("expo::", "").replace("tma::", "")
  m = re.match(r"(\w+)<(.*)>$", name)
  if m and "at::" not in name:
    args = re.sub(r"\(int\)|\(bool\)", "", m.group(2))
    return "%s<%s>" % (m.group(1), args[:40])
  if "at::" in name or "at_cuda" in name:
    k = re.search(r"(\w+_kernel\w*)", name)
    return "torch:" + (k.group(1) if k else name[:40])
  return name[:60]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("--iters", type=int, default=5)
  ap.add_argument("--batch", type=int, default=64)
  ap.add_argument("--out", default="gpurun_out/timeline.json")
  ap.add_argument("--host-replay", action="store_true")
  args = ap.parse_args()
  import torch
  import torch.distributed as dist
  from torch.profiler import ProfilerActivity, profile
  from exposure_b200.replay import DeviceReplayMemory, ReplayMemory, SyntheticProvider
  from exposure_b200.trainer import Trainer, default_cfg
  world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
  torch.cuda.set_device(local)
  dev = torch.device("cuda", local)
  if world > 1:
    dist.init_process_group("nccl", device_id=dev)
  cfg = default_cfg()
  cfg.batch_size = args.batch
  cfg.replay_memory_size = 2 * args.batch
  t = Trainer(cfg, dev, seed=0)
  Mem = ReplayMemory if args.host_replay else DeviceReplayMemory
  mem = Mem(cfg, SyntheticProvider(dev, "raw", 100 + rank), SyntheticProvider(dev, "real", 200 + rank), dev, seed=rank)
  t.attach_memory(mem, torch.Generator(device=dev).manual_seed(300 + rank))
  t.train_iteration(0, giters=2 * cfg.test_steps + 2, citers=1)
  t.enable_graphs(args.batch)
  if not args.host_replay and t._graph_apply:
    t.enable_iteration_graph()
  it = 1
  for _ in range(5):
    t.train_iteration(it, giters=1, citers=5)
    it += 1
  torch.cuda.synchronize()
  with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(args.iters):
      t.train_iteration(it, giters=1, citers=5)
      it += 1
    torch.cuda.synchronize()
  if world > 1:
    dist.barrier()
  if rank != 0:
    dist.destroy_process_group()
    return
  evs = []
  for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start:
      evs.append((e.time_range.start, e.time_range.end, e.name))
  evs.sort()
  if not evs:
    print("no CUDA events captured")
    return
  t0, t1 = evs[0][0], max(e[1] for e in evs)
  span = t1 - t0
  # union of intervals, gaps, concurrency
  busy = 0.0
  gaps = []
  cur_s, cur_e = evs[0][0], evs[0][1]
  for s, e, _ in evs[1:]:
    if s > cur_e:
      busy += cur_e - cur_s
      gaps.append(s - cur_e)
      cur_s, cur_e = s, e
    else:
      cur_e = max(cur_e, e)
  busy += cur_e - cur_s
  ksum = sum(e - s for s, e, _ in evs)
  per = collections.OrderedDict()
  for s, e, n in evs:
    d = per.setdefault(short(n), [0, 0.0])
    d[0] += 1
    d[1] += e - s
  # exclusive time: for every instant, 1/n of it is charged to each of the n kernels running then, and
  # `solo` is the time a kernel runs with nothing beside it (a proxy for its share of the critical path)
  pts = []
  for i, (s, e, n) in enumerate(evs):
    pts.append((s, 1, i))
    pts.append((e, 0, i))
  pts.sort()
  active = set()
  share = collections.defaultdict(float)
  solo = collections.defaultdict(float)
  prev = pts[0][0]
  for tm, kind, i in pts:
    if tm > prev and active:
      w = (tm - prev) / len(active)
      for j in active:
        share[short(evs[j][2])] += w
      if len(active) == 1:
        solo[short(evs[next(iter(active))][2])] += tm - prev
    prev = tm
    if kind:
      active.add(i)
    else:
      active.discard(i)
  gaps.sort()
  big = [g for g in gaps if g > 20.0]
  res = {
      "iters": args.iters, "batch": args.batch, "pdl": os.environ.get("EXPOSURE_PDL", "default"), "n_gpus": world,
      "replay": "host" if args.host_replay else "device, whole iteration one graph",
      "transport": "peer-memory all-reduce + Adam kernel" if t._peer is not None else ("dist.all_reduce" if world > 1 else "single GPU"),
      "span_us_per_iter": span / args.iters, "busy_us_per_iter": busy / args.iters,
      "idle_us_per_iter": (span - busy) / args.iters, "kernel_sum_us_per_iter": ksum / args.iters,
      "avg_concurrency_when_busy": ksum / busy, "kernels_per_iter": len(evs) / args.iters,
      "gaps": {"count_per_iter": len(gaps) / args.iters, "median_us": gaps[len(gaps) // 2] if gaps else 0,
               "p90_us": gaps[int(len(gaps) * 0.9)] if gaps else 0, "sum_us_per_iter": sum(gaps) / args.iters,
               "over_20us_count_per_iter": len(big) / args.iters, "over_20us_sum_per_iter": sum(big) / args.iters},
      "top_kernels": [{"kernel": k, "launches_per_iter": v[0] / args.iters, "us_per_iter": v[1] / args.iters,
                       "avg_us": v[1] / v[0], "share_us_per_iter": share[k] / args.iters,
                       "solo_us_per_iter": solo[k] / args.iters}
                      for k, v in sorted(per.items(), key=lambda kv: -share[kv[0]])[:45]],
  }
  os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
  json.dump(res, open(args.out, "w"), indent=1)
  print(json.dumps({k: v for k, v in res.items() if k != "top_kernels"}, indent=1))
  print("%-58s %7s %8s %9s %9s %9s" % ("kernel", "n/iter", "avg us", "sum/iter", "share", "solo"))
  for k in res["top_kernels"][:40]:
    print("%-58s %7.1f %8.2f %9.1f %9.1f %9.1f" % (k["kernel"][:58], k["launches_per_iter"], k["avg_us"], k["us_per_iter"],
                                                    k["share_us_per_iter"], k["solo_us_per_iter"]))
  if world > 1:
    dist.destroy_process_group()


if __name__ == "__main__":
  main()
