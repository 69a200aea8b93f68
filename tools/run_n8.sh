#!/bin/bash
# One 8-GPU box call: data-parallel check, train bench at N = 8 / 4 (peer-memory transport) and N = 8 with NCCL outside the
# graph (A/B), the 4K sweep at N = 8 / 4, and a CUPTI timeline of rank 0 at N = 8.  Outputs under gpurun_out/.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 240 $TR --nproc-per-node 8 --master-port 29701 tools/dp_check.py 2>&1 | grep dp_check > gpurun_out/r2_dp_check_n8.log
tail -3 gpurun_out/r2_dp_check_n8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --steps 20 --warmup 5 --roofline-batch 0 > gpurun_out/r2_bench_train_n8.json 2> gpurun_out/r2_bench_train_n8.err
tail -2 gpurun_out/r2_bench_train_n8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29703 bench.py --gpus 4 --steps 20 --warmup 5 --roofline-batch 0 > gpurun_out/r2_bench_train_n4.json 2>/dev/null
EXPOSURE_DP_TRANSPORT=allreduce timeout 300 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --steps 20 --warmup 5 --roofline-batch 0 > gpurun_out/r2_bench_train_n8_nccl.json 2>/dev/null
timeout 200 $TR --nproc-per-node 8 --master-port 29706 tools/sweep_4k.py 2>/dev/null | grep workload > gpurun_out/r2_4k_sweep_n8.jsonl
timeout 200 $TR --nproc-per-node 4 --master-port 29707 tools/sweep_4k.py 2>/dev/null | grep workload > gpurun_out/r2_4k_sweep_n4.jsonl
timeout 240 $TR --nproc-per-node 8 --master-port 29708 tools/train_timeline.py --out gpurun_out/r2_timeline_n8.json > gpurun_out/r2_timeline_n8.txt 2>&1
for f in gpurun_out/r2_bench_train_n8.json gpurun_out/r2_bench_train_n4.json gpurun_out/r2_bench_train_n8_nccl.json; do
  python -c "import json; d=json.load(open('$f')); print('$f', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
done
cut -c60-260 gpurun_out/r2_4k_sweep_n8.jsonl
grep -E "span_us|busy_us|idle_us|dp_allreduce" gpurun_out/r2_timeline_n8.txt
