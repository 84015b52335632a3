#!/bin/bash
# 2-GPU validation: same fit on 1 and 2 ranks, then the bench line at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/dist2_gpus.txt 2>&1
( timeout 300 python tools/dist_check.py --single 2>&1 | tail -3 ) > gpurun_out/dist_single.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | tail -8 ) > gpurun_out/dist_check2.log
( timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/bench_dist_n1.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 --no-cpu-baseline 2>&1 | tail -3 ) > gpurun_out/bench_dist_n2.log
cat gpurun_out/dist_single.log gpurun_out/dist_check2.log; for f in gpurun_out/bench_dist_n1.log gpurun_out/bench_dist_n2.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; grep -o '"e2e": {[^}]*}' $f; tail -c 300 $f; done
