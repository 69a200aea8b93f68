#!/bin/bash
# single-GPU check of the round-2 late kernels: small-Cin dgrad (smem-staged), FC GEMMs with 4 K-steps per round trip,
# forward-only chain kernel with parallel set-up
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_nn_gpu.py tests/test_train_step_gpu.py tests/test_train_glue_gpu.py tests/test_eval_path_gpu.py tests/test_nn_tma_gpu.py tests/test_filter_api_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 20 --warmup 5 --roofline-batch 0 > gpurun_out/r2m_bench_train_n1.json 2> gpurun_out/r2m_err.log; tail -2 gpurun_out/r2m_err.log
timeout 300 python bench.py --workload eval --batch 8 --height 2160 --width 3840 --steps 20 --warmup 3 > gpurun_out/r2m_bench_eval_4k.json 2>> gpurun_out/r2m_err.log
timeout 300 python bench.py --workload eval --batch 256 --steps 50 --warmup 5 > gpurun_out/r2m_bench_eval_b256.json 2>> gpurun_out/r2m_err.log
timeout 300 python bench.py --workload chain8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_chain8_pdl0.json 2>> gpurun_out/r2m_err.log
EXPOSURE_PDL=1 timeout 300 python bench.py --workload chain8 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r2m_bench_chain8_pdl1.json 2>> gpurun_out/r2m_err.log
python - <<'PY'
import json
for f in ("train_n1","eval_4k","eval_b256","chain8_pdl0","chain8_pdl1"):
  try:
    d=json.load(open("gpurun_out/r2m_bench_%s.json"%f)); r=d.get("roofline",{})
    print(f, round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "frac", r.get("frac"), "per_step", r.get("per_step_frac"), r.get("filters_applied"))
  except Exception as e: print(f, "ERR", e)
PY
