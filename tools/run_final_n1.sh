#!/bin/bash
# Final single-GPU evidence of the round: GPU test suite, smoke, the bench lines exactly as the driver runs them, the ncu
# launch list of the train command, a CUPTI timeline and the isolated op times.  Outputs under gpurun_out/r2final_*.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/r2final_gpu_tests.log; cat gpurun_out/r2final_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final_bench_train_n1.json 2> gpurun_out/r2final_err.log
timeout 300 python bench.py --workload chain8 --steps 50 --warmup 5 > gpurun_out/r2final_bench_chain8_n1.json 2>> gpurun_out/r2final_err.log
timeout 300 python bench.py --workload eval --batch 8 --height 2160 --width 3840 --steps 20 --warmup 3 > gpurun_out/r2final_bench_eval_4k.json 2>> gpurun_out/r2final_err.log
timeout 300 python bench.py --workload eval --batch 256 --steps 50 --warmup 5 > gpurun_out/r2final_bench_eval_b256.json 2>> gpurun_out/r2final_err.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r2final_launches_train.csv python bench.py --steps 2 --warmup 1 --roofline-batch 0 --no-cpu-baseline > /dev/null 2>&1
timeout 300 python tools/train_timeline.py --out gpurun_out/r2final_timeline_n1.json > gpurun_out/r2final_timeline_n1.txt 2>&1
timeout 300 python tools/op_bench.py > gpurun_out/r2final_op_bench.txt 2>&1
python - <<'PY'
import json
for f in ("train_n1", "chain8_n1", "eval_4k", "eval_b256"):
  try:
    d = json.load(open("gpurun_out/r2final_bench_%s.json" % f)); r = d.get("roofline") or {}
    print(f, round(d["value"], 1), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "frac", r.get("frac"), "per_step",
          r.get("per_step_frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
  except Exception as e:
    print(f, "ERR", e)
PY
