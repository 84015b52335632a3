#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/dist_check.py --single 2>&1 | tail -1 ) > gpurun_out/dist_single.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | tail -4 ) > gpurun_out/dist_check2.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/bench_dist_n2.log
cat gpurun_out/dist_single.log gpurun_out/dist_check2.log; for f in gpurun_out/bench_dist_n2.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f; done
