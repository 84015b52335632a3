#!/bin/bash
# 8 GPUs: cfg5 (N = 1M, 16 inputs, 3-way): the row- and candidate-sharded fit must reproduce the 1-GPU fit; bench lines
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2k_gpus.txt 2>&1
( timeout 200 python tools/dist_check.py --single --cfg cfg5 --rows 200000 2>&1 | tail -1 ) > gpurun_out/r2k_single.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py --cfg cfg5 --rows 200000 2>&1 | tail -3 ) > gpurun_out/r2k_check.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload cfg5 --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2k_bench_cfg5_n8.log
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2k_bench_cfg4_n8.log
cat gpurun_out/r2k_single.log gpurun_out/r2k_check.log; for f in gpurun_out/r2k_bench_cfg5_n8.log gpurun_out/r2k_bench_cfg4_n8.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; done
