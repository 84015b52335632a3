#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_candidates.py tests/test_gpu_fit.py -m gpu -x -q --durations=8 -k "blocked or whole_device or cfg5" 2>&1 | tail -25 ) > gpurun_out/r2i_pytest.log; cat gpurun_out/r2i_pytest.log
( timeout 600 python bench.py --workload cfg5 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -3 ) > gpurun_out/r2i_bench_cfg5.log; cat gpurun_out/r2i_bench_cfg5.log
