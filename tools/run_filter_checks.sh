#!/bin/bash
# closed-form S+ forward, Contrast forward with one reciprocal: parity suites that touch the filters + the benches it moves
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_filters_gpu.py tests/test_reference_golden_gpu.py tests/test_chain_fused_gpu.py tests/test_eval_path_gpu.py tests/test_filter_api_gpu.py tests/test_baseline_shapes_gpu.py -m gpu -q 2>&1 | tail -6
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 300 python bench.py --workload chain8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2ah_bench_chain8_n1.json 2>/dev/null
timeout 300 python bench.py --workload chain8 --batch 256 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2ah_bench_chain8_b256.json 2>/dev/null
python - <<'PY'
import json
for f in ("chain8_n1", "chain8_b256"):
  try:
    d = json.load(open("gpurun_out/r2ah_bench_%s.json" % f)); r = d["roofline"]
    print(f, round(d["value"], 1), round(d["ms_per_step"], 4), r.get("frac"), r.get("per_step_frac"), r.get("per_step_worst_kernel"),
          r.get("per_step_worst_kernel_frac"), r.get("k_filter_fwd_satplus_frac"))
  except Exception as e:
    print(f, "ERR", e)
PY
