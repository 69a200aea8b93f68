#!/usr/bin/env python
"""Turn ncu artefacts brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python profiles/summarize.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python profiles/summarize.py full gpurun_out/prof_r1_filter.ncu-rep profiles/r1_filter_ncu.md
"""
import collections
import csv
import re
import subprocess
import sys


def short(name):
  name = re.sub(r"^void ", "", name)
  m = re.match(r"(expo::)?(\w+)<([^>]*)>", name)
  if m and ("filter_step" in name or "gemm_kernel" in name or "filter_chain" in name):
    return "%s<%s>" % (m.group(2), m.group(3))
  if "at::" in name or "templates::" in name:
    k = re.search(r"(\w+_kernel\w*)", name)
    return "torch:" + (k.group(1) if k else name[:40])
  return name.split("(")[0][:60]


def launches(src, dst):
  rows = [r for r in csv.reader(open(src)) if len(r) > 5 and r[0].isdigit()]
  tot = collections.OrderedDict()
  for r in rows:
    d = tot.setdefault(short(r[4]), [0, 0.0])
    d[0] += 1
    d[1] += float(r[-1])
  total = sum(v[1] for v in tot.values())
  with open(dst, "w") as fh:
    fh.write("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache + serialised:\n"
             "# compare SHARES, not absolutes).  Source: %s, %d launches, %.3f ms total\n\n" % (src, len(rows), total / 1e6))
    fh.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
      fh.write("| `%s` | %d | %.1f | %.2f | %.1f%% |\n" % (k, n, t / 1e3, t / n / 1e3, 100 * t / total))
  print(open(dst).read())


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "smsp__inst_executed.sum", "sm__inst_executed_pipe_xu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]


def full(src, dst, note=""):
  out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}
  with open(dst, "w") as fh:
    fh.write("# ncu --set full --clock-control none capture (%s)\n%s\n" % (src, note))
    for r in rows[2:]:
      fh.write("\n## `%s`\n\n| metric | value | unit |\n|---|---:|---|\n" % short(r[idx["Kernel Name"]]))
      for w in WANT:
        if w in idx:
          fh.write("| %s | %s | %s |\n" % (w, r[idx[w]], units[idx[w]]))
      t = float(r[idx["gpu__time_duration.sum"]])
      rd = float(r[idx["dram__bytes_read.sum"]]); wr = float(r[idx["dram__bytes_write.sum"]])
      fh.write("\nDRAM traffic = %.1f MB read + %.1f MB written per launch (units as reported above).\n" % (rd, wr))
  print(open(dst).read()[:3000])


UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME_US = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}
FNAMES = ("exposure", "gamma", "wb", "satplus", "tone", "contrast", "wnb", "color", "level", "vignet")


def traffic(src, dst, config):
  """DRAM bytes per launch of the filter kernels in an `ncu --set full` report -> the small JSON bench.py reads to
  fill roofline.traffic (keys = bench.py's kernel names).  One entry per captured configuration
  ({"captures": [{"config", "source", "kernels"}, ...]}); a capture of an already known configuration is merged."""
  import json
  raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(raw.splitlines()))
  hdr, units = rows[0], rows[1]
  idx = {h: i for i, h in enumerate(hdr)}
  try:
    doc = json.load(open(dst))
    caps = doc["captures"] if "captures" in doc else [doc]
  except (OSError, ValueError, KeyError):
    caps = []
  ent = next((c for c in caps if c.get("config") == config), None)
  if ent is None:
    ent = {"config": config, "source": src, "kernels": {}}
    caps.append(ent)
  elif src not in ent["source"]:
    ent["source"] = "%s + %s" % (ent["source"], src)
  res = ent["kernels"]

  def put(name, r):
    rd = float(r[idx["dram__bytes_read.sum"]]) * UNIT[units[idx["dram__bytes_read.sum"]]]
    wr = float(r[idx["dram__bytes_write.sum"]]) * UNIT[units[idx["dram__bytes_write.sum"]]]
    res[name] = {"dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
                 "duration_us_under_ncu": float(r[idx["gpu__time_duration.sum"]]) * TIME_US[units[idx["gpu__time_duration.sum"]]],
                 "kernel": short(r[idx["Kernel Name"]])}

  for r in rows[2:]:
    kn = r[idx["Kernel Name"]]
    if "filter_chain_static_kernel" in kn or "filter_chain_fwd_bwd_kernel" in kn:
      put("filter_chain_fwd_bwd", r)          # bench.py's name of the whole-chain launch, whichever kernel serves it
      continue
    m = re.search(r"filter_step_tma_kernel<(?:\(int\))?(\d+), (?:\(bool\))?(\w+), (?:\(bool\))?(\w+)", kn)
    if not m:
      continue
    fid, bwd, gx = int(m.group(1)), m.group(2) in ("1", "true"), m.group(3) in ("1", "true")
    put(("filter_bwd_" if bwd and gx else "filter_bwd_paramonly_" if bwd else "filter_fwd_") + FNAMES[fid], r)
  json.dump({"note": "ncu --set full --clock-control none, one launch each; writes still in L2 at kernel end are not "
                     "counted by dram__bytes_write", "captures": caps}, open(dst, "w"), indent=1, sort_keys=True)
  print(open(dst).read()[:1500])


if __name__ == "__main__":
  {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](*sys.argv[2:])
