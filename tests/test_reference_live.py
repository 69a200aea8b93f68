"""CPU, build container only: the oracle against the reference's own Python on RANDOM cases (60 random
shapes / seeds / logit scales over all ten filters, masked and unmasked), run live in a subprocess
(tests/golden/live_check.py) -- the fixed vectors of reference_golden.npz are one draw, this is sixty."""
import json
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs /root/reference (build container only)")


def test_oracle_matches_reference_code_on_random_cases():
  r = subprocess.run([sys.executable, os.path.join(HERE, "golden", "live_check.py"), "60"], capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-3000:]
  d = json.loads(r.stdout.strip().splitlines()[-1])
  assert d["cases"] == 60
  w = d["worst"]
  # fp64 vs fp64; the bound is float32 rounding of constants (see tests/test_reference_golden.py)
  assert w["param"] < 5e-7 and w["y"] < 5e-6 and w["glogits"] < 2e-5 and w["gx"] < 2e-5, w
  assert w["mask"] < 5e-6 and w["masked"] < 5e-6 and w["stats"] < 5e-6, w
