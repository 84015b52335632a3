#!/bin/bash
# session 4, call d: K2 tiles of equal work; correctness + sweep + cfg4 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gram.py -x -q 2>&1 | tail -4 > gpurun_out/r4d_pytest_gram.log
cat gpurun_out/r4d_pytest_gram.log
timeout 900 python tools/gram_sweep.py --reps 6 --variants 'auto;FOKL_GRAM_WARPS=12' > gpurun_out/r4d_sweep.txt 2>&1
cat gpurun_out/r4d_sweep.txt
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4d_bench_cfg4.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4d_bench_cfg4.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4d_bench_cfg4.log
