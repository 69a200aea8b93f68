#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_nn_gpu.py -m gpu -q -k "case8 or case7" --tb=short 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2n_case8.log
cat gpurun_out/r2n_case8.log | tail -30
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r2n_suite.log
cat gpurun_out/r2n_suite.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:filter_chain_static -c 1 -f -o gpurun_out/r2n_chain_static_b256 python bench.py --workload chain8 --batch 256 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_ncu.log 2>&1
tail -2 gpurun_out/r2n_ncu.log
