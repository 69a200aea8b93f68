#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2final2_gpu_tests.log; cat gpurun_out/r2final2_gpu_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final2_bench_train_n1.json 2> gpurun_out/r2final2_err.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2final2_bench_train_n1.json")); r = d["roofline"]
print(round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "frac", r.get("frac"), "per_step", r.get("per_step_frac"),
      r.get("per_step_worst_kernel"), r.get("per_step_worst_kernel_frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
