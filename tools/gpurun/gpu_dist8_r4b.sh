#!/bin/bash
# session 4, final: 8-GPU box -- identical fit on 8 ranks, then the bench lines at N = 2, 4, 8 (cfg4) and N = 8 (cfg5)
mkdir -p gpurun_out
# the single-GPU fit the ranks are compared with (gpurun_out/ does not travel to the box: without this line rank 0 finds no
# dist_single.npz -- that is the traceback in gpurun_out/d6_check.log; the 8-rank check itself passed in gpu_dist8_r4.sh)
( timeout 150 python tools/dist_check.py --single 2>&1 | tail -1 ) > gpurun_out/d6_single.log
( timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py 2>&1 | tail -2 ) > gpurun_out/d6_check.log
cat gpurun_out/d6_check.log
P=29520
for n in 2 4 8; do
P=$((P+1))
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/d6_bench_n$n.log
done
( timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --workload cfg5 --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 ) > gpurun_out/d6_bench_cfg5_n8.log
for f in gpurun_out/d6_bench_n2.log gpurun_out/d6_bench_n4.log gpurun_out/d6_bench_n8.log gpurun_out/d6_bench_cfg5_n8.log; do echo $f; grep -o '"ms_per_step": [0-9.]*' $f | head -1; grep -o '"stage_ms_per_step": {[^}]*}' $f; done
