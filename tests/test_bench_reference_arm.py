"""CPU: `bench.py --impl reference` (the reference's CPU path = the oracle port timed on the host cores)
prints ONE JSON line with the contract's keys; under a multi-rank launch only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None, extra=()):
  env = dict(os.environ, **(env_extra or {}))
  return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                         "--size", "64", "--batch", "4"] + list(extra), capture_output=True, text=True, env=env, timeout=300)


import pytest


@pytest.mark.parametrize("workload", ["train", "chain8", "eval"])
def test_reference_arm_json_line(workload):
  """The default workload is the train step (the config BASELINE.json's metric is quoted on); every workload's
  reference arm runs the FULL per-GPU batch of the native arm's config."""
  r = _run(extra=() if workload == "train" else ("--workload", workload))
  assert r.returncode == 0, r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if l.strip()]
  assert len(lines) == 1
  d = json.loads(lines[0])
  assert d["impl"] == "reference" and d["metric"] == "images/sec" and d["unit"] == "images/s"
  assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
  assert d["config"]["workload"].startswith(workload) and d["dtype"] == "f32" and d["vs_baseline"] is None
  assert d["config"]["batch_per_gpu"] == 4 and "full per-GPU batch" in d["cpu_baseline"]["sample"]
  cb = d["cpu_baseline"]
  assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
  assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
  assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
  r = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
  assert r.returncode == 0 and r.stdout.strip() == ""
