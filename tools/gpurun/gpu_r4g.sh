#!/bin/bash
# session 4, call g: cluster Jacobi without the confirming sweep + branch-free rsqrt; candidates tests + cfg4 bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_candidates.py -x -q 2>&1 | tail -4 > gpurun_out/r4g_pytest_cand.log
cat gpurun_out/r4g_pytest_cand.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r4g_bench_cfg4.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r4g_bench_cfg4.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r4g_bench_cfg4.log
timeout 300 python tools/eig_diag.py 220 2>&1 | tail -5
