#!/bin/bash
# session 4: 2-GPU check of the reordered loop + control communicator
mkdir -p gpurun_out
( timeout 200 python tools/dist_check.py --single 2>&1 | tail -1 ) > gpurun_out/d5_single.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dist_check.py 2>&1 | tail -2 ) > gpurun_out/d5_check.log
cat gpurun_out/d5_single.log gpurun_out/d5_check.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/d5_bench_n2.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/d5_bench_n2.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/d5_bench_n2.log
