#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_fit.py tests/test_gpu_candidates.py -m gpu -x -q -k "nested or cfg5 or cfg4_shaped" 2>&1 | tail -6 ) > gpurun_out/r2r_pytest.log; cat gpurun_out/r2r_pytest.log
( timeout 600 python bench.py --workload cfg5 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2r_bench_cfg5.log
echo cfg5; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2r_bench_cfg5.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2r_bench_cfg5.log; grep -o '"eig_solves[^}]*' gpurun_out/r2r_bench_cfg5.log
