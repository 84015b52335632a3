#!/bin/bash
# cfg5 at N GPUs (N from the environment): bench line
mkdir -p gpurun_out
N=${NGPU:-2}
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --workload cfg5 --gpus $N --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/r2u_bench_cfg5_n$N.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2u_bench_cfg5_n$N.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2u_bench_cfg5_n$N.log
