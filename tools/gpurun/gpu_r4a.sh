#!/bin/bash
# session 4, call a: loose-block K2 (correctness, sweep, cfg4 bench) + host-side profile of one cfg4 fit
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gram.py -x -q 2>&1 | tail -4 > gpurun_out/r4a_pytest_gram.log
cat gpurun_out/r4a_pytest_gram.log
timeout 600 python tools/gram_sweep.py --reps 6 --variants 'auto;FOKL_GRAM_WARPS=15;FOKL_GRAM_WARPS=12' > gpurun_out/r4a_sweep.txt 2>&1
cat gpurun_out/r4a_sweep.txt
timeout 600 python tools/host_profile.py > gpurun_out/r4a_host_profile.txt 2>&1
head -40 gpurun_out/r4a_host_profile.txt
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4a_bench_cfg4.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4a_bench_cfg4.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4a_bench_cfg4.log
