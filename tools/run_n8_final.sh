#!/bin/bash
# One 8-GPU box call on the final tree: replica check + exchange-kernel timing, the train bench exactly as the driver
# launches it at N = 8 and N = 4.  Outputs under gpurun_out/.
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
timeout 200 $TR --nproc-per-node 8 --master-port 29711 tools/dp_check.py 2>&1 | grep dp_check > gpurun_out/r2final_dp_check_n8.log
tail -6 gpurun_out/r2final_dp_check_n8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29728 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2final_bench_train_n8.json 2> gpurun_out/r2final_bench_train_n8.err
timeout 300 $TR --nproc-per-node 4 --master-port 29724 bench.py --gpus 4 --steps 20 --warmup 5 --roofline-batch 0 > gpurun_out/r2final_bench_train_n4.json 2> gpurun_out/r2final_bench_train_n4.err
for f in gpurun_out/r2final_bench_train_n8.json gpurun_out/r2final_bench_train_n4.json; do
  python -c "import json; d=json.load(open('$f')); print('$f', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],4), round(d['e2e']['value'],1))"
done
