#!/bin/bash
# session 4: 8-GPU box -- identical fit on 8 ranks, then the bench lines at N = 1, 2, 4, 8 (cfg4) and N = 8 (cfg5)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/d4_gpus.txt 2>&1
( timeout 200 python tools/dist_check.py --single 2>&1 | tail -1 ) > gpurun_out/d4_single.log
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | tail -2 ) > gpurun_out/d4_check.log
cat gpurun_out/d4_single.log gpurun_out/d4_check.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/d4_bench_n1.log
P=29520
for n in 2 4 8; do
P=$((P+1))
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/d4_bench_n$n.log
done
( timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --workload cfg5 --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/d4_bench_cfg5_n8.log
for f in gpurun_out/d4_bench_n1.log gpurun_out/d4_bench_n2.log gpurun_out/d4_bench_n4.log gpurun_out/d4_bench_n8.log gpurun_out/d4_bench_cfg5_n8.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; done
