#!/bin/bash
mkdir -p gpurun_out
for nb in 26 64; do echo "FOKL_SECULAR_BISECT=$nb"; for a in "40 12 gauss" "400 100 gauss" "400 100 spline" "2072 100 gauss"; do FOKL_SECULAR_BISECT=$nb timeout 300 python tools/nested_check.py $a 2>&1 | grep -v Warn | head -1; done; done > gpurun_out/r2y_nested.log; cat gpurun_out/r2y_nested.log
( timeout 600 python -m pytest tests/test_gpu_candidates.py tests/test_gpu_fit.py -m gpu -x -q -k "nested or cfg5" 2>&1 | tail -4 ) > gpurun_out/r2y_pytest.log; cat gpurun_out/r2y_pytest.log
( timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2y_bench_cfg5.log
echo cfg5; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2y_bench_cfg5.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2y_bench_cfg5.log
