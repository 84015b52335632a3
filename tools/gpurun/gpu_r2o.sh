#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_candidates.py tests/test_gpu_fit.py tests/test_gpu_fullsize.py -m gpu -x -q --durations=5 2>&1 | tail -25 ) > gpurun_out/r2o_pytest.log; cat gpurun_out/r2o_pytest.log
( timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2o_bench_cfg4.log
echo cfg4; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2o_bench_cfg4.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2o_bench_cfg4.log; grep -o '"work_per_step": {[^}]*}' gpurun_out/r2o_bench_cfg4.log | grep -o '"eig_solves.*'
( timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 ) > gpurun_out/r2o_bench_cfg5.log
echo cfg5; grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2o_bench_cfg5.log | head -1; grep -o '"stage_ms_per_step": {[^}]*}' gpurun_out/r2o_bench_cfg5.log; grep -o '"work_per_step": {[^}]*}' gpurun_out/r2o_bench_cfg5.log | grep -o '"eig_solves.*'
