#!/bin/bash
# 8-GPU validation: the row-sharded fit must reproduce the 1-GPU fit; then the bench line at N = 8
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/dist8_gpus.txt 2>&1
( timeout 200 python tools/dist_check.py --single 2>&1 | tail -1 ) > gpurun_out/dist8_single.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | tail -2 ) > gpurun_out/dist8_check.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_dist_n8.log
cat gpurun_out/dist8_single.log gpurun_out/dist8_check.log; for f in gpurun_out/bench_dist_n8.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f; done
